"""Sharded (multi-GPU) prove, exercised on ONE device through the local communicator:
all ranks' stages run in lockstep and every collective is a device copy, so the sharded
code path (coset-range LDE, subtree roots + replicated top, owner-computed quotient chunks,
sharded FRI rounds, owner-written query openings) is checked bit-for-bit against the
single-GPU proof and the oracle.  With >= 2 GPUs the same test also runs over NCCL
(tests/run_sharded_nccl.py under torchrun)."""
import numpy as np
import pytest

from oracle import air as OA
from oracle import field as F
from oracle import stark as OS
from oracle import trace as OT

pytestmark = pytest.mark.gpu


def _instance(log_n, c, seed):
    rng = F.SplitMix64(seed)
    alpha, delta = rng.next_fr(), rng.next_fr()
    cfgs, trace = OT.build_trace([OT.synthetic_permutation_input(seed, c, 1 << log_n)], alpha, delta)
    return cfgs, trace, [alpha, delta]


def _gpu_cfgs(pkg, cfgs):
    return [pkg.AirPermutationConfig(c.a_columns_ids, c.b_columns_ids, c.b_inverse_id, c.check_id) for c in cfgs]


@pytest.mark.parametrize("log_n,c,world,blowup", [(4, 2, 1, 3), (4, 2, 2, 3), (5, 3, 4, 3), (6, 1, 8, 3), (5, 2, 2, 1),
                                                  (10, 3, 2, 3), (11, 2, 4, 3), (12, 1, 8, 3),
                                                  # more ranks than cosets (BASELINE configs[3]: blowup 2 and 4 on 8 GPUs): a rank owns a
                                                  # fraction of a coset, evaluates the NEXT rows itself, chunks assembled by a sum
                                                  (4, 2, 4, 1), (5, 1, 8, 1), (6, 3, 8, 2), (6, 2, 4, 1), (10, 3, 8, 1), (12, 2, 8, 2),
                                                  (13, 1, 16, 1)])
def test_sharded_prove_equals_single_gpu(pkg, gctx, p2params, log_n, c, world, blowup):
    cfgs, trace, publics = _instance(log_n, c, 100 + log_n + world)
    fri = pkg.FriConfig(log_blowup=blowup, num_queries=17)
    g = _gpu_cfgs(pkg, cfgs)
    single = pkg.prove(gctx, fri, g, trace, publics)
    comm = pkg.Comm.local(gctx, world)
    sharded = pkg.prove_sharded(comm, fri, g, trace, publics)
    comm.close()
    assert np.array_equal(single.words, sharded.words)
    if log_n <= 6:  # and both equal the oracle's proof, which its verifier accepts
        ofri = OS.FriConfig(log_blowup=blowup, num_queries=17)
        oproof = OS.prove(p2params, ofri, cfgs, trace, publics)
        gd, _ = sharded.to_dict()
        assert gd == oproof
        OS.verify(p2params, ofri, cfgs, gd, publics)


def test_sharded_rejects_too_many_ranks(pkg, gctx):
    """A rank may own a fraction of a coset, but not fewer than 8 rows of it (the alignment of the 1/(x - z) tables)."""
    cfgs, trace, publics = _instance(3, 1, 5)
    comm = pkg.Comm.local(gctx, 8)
    with pytest.raises(pkg.BackendError, match="too many"):
        pkg.prove_sharded(comm, pkg.FriConfig(log_blowup=1), _gpu_cfgs(pkg, cfgs), trace, publics)  # 8 ranks, 2 cosets of 8 rows
    comm.close()


def test_sharded_nccl_two_gpus(pkg):
    """Real NCCL path on every device of the box (2, 4 or 8 ranks; the 1-GPU box covers the same code via the local
    communicator): permutation AIR, an odd-width lookup AIR, and a blowup-2 case that has more ranks than cosets from 4 up."""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    from pathlib import Path
    script = Path(__file__).parent / "run_sharded_nccl.py"
    n = torch.cuda.device_count()
    n = 8 if n >= 8 else 4 if n >= 4 else 2
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr",
                        "127.0.0.1", "--master-port", "29517", str(script)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "sharded nccl ok" in r.stdout
