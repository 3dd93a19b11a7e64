#!/bin/bash
# Re-measurement of single points of the multi-GPU sweep (see tools/scale_sweep.sh) on an 8-GPU box, after the NCCL parity test.
TAG=$1; OUT=gpurun_out
run() { local n=$1 name=$2; shift 2
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) bench.py --gpus $n "$@" > $OUT/${TAG}_${name}_n$n.json 2> $OUT/${TAG}_${name}_n$n.err
  echo "$name n=$n rc=$? $(python -c "
import json
d=json.loads(open('$OUT/${TAG}_${name}_n$n.json').read().strip().splitlines()[-1])
print(round(d['value']*1e3,2),'ms e2e',round(d['e2e']['value']*1e3,2),d['stages_ms'],'eq',d.get('sharded_equals_single'),'cfg3',(d.get('extra') or {}).get('cfg3',{}).get('value'))")"; }
python -m pytest tests/test_gpu_sharded.py -m gpu -x -q -k nccl > $OUT/${TAG}_pytest_nccl.log 2>&1; echo "nccl pytest rc=$?"; tail -2 $OUT/${TAG}_pytest_nccl.log
run 8 cfg1 --steps 5 --warmup 3 --no-cpu-baseline
run 8 cfg4b --steps 5 --warmup 3 --cols 32 --log-n 20 --log-blowup 2 --cfg3 0 --no-cpu-baseline
