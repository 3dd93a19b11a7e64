#!/bin/bash
# Multi-GPU session on ONE box (gpurun --gpus 8): the NCCL tests, then bench.py at N = 1, 2, 4, 8 for the headline workload
# (cfg-1, with cfg-3 under extra.cfg3) and for BASELINE configs[3] (3x32 columns, 2^20 rows, blowup 2 and 4), which needs
# more ranks than cosets at N = 4 / 8.  Results: gpurun_out/<tag>_*.json
TAG=${1:-sweep}
OUT=gpurun_out
run() {  # n, name, extra args...
  local n=$1 name=$2; shift 2
  if [ "$n" = 1 ]; then python bench.py --gpus 1 "$@" > $OUT/${TAG}_${name}_n1.json 2> $OUT/${TAG}_${name}_n1.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n "$@" > $OUT/${TAG}_${name}_n$n.json 2> $OUT/${TAG}_${name}_n$n.err; fi
  echo "$name n=$n rc=$? $(python -c "
import json,sys
try:
    d=json.loads(open('$OUT/${TAG}_${name}_n$n.json').read().strip().splitlines()[-1])
    print(round(d['value']*1e3,2),'ms e2e',round(d['e2e']['value']*1e3,2),'fri',d['stages_ms'].get('fri_commit_phase'),'eq',d.get('sharded_equals_single'), 'cfg3', (d.get('extra') or {}).get('cfg3',{}).get('value'))
except Exception as e: print('ERR',e)")"
}
LSP_SKIP_BIG=1 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_main_driver.py -m gpu -x -q > $OUT/${TAG}_pytest_nccl.log 2>&1; echo "nccl pytest rc=$?"; tail -3 $OUT/${TAG}_pytest_nccl.log
for n in 8 4 2 1; do run $n cfg1 --steps 5 --warmup 3 --no-cpu-baseline; done
for n in 8 4 2 1; do run $n cfg4a --steps 3 --warmup 3 --cols 32 --log-n 20 --log-blowup 1 --cfg3 0 --no-cpu-baseline; done
for n in 8 4 2 1; do run $n cfg4b --steps 3 --warmup 3 --cols 32 --log-n 20 --log-blowup 2 --cfg3 0 --no-cpu-baseline; done
