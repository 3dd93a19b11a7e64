#!/usr/bin/env python
"""One prove of the bench workload, no timing -- the command line ncu wraps
(`ncu ... python tools/profile_prove.py --log-n 19`).  Numbers printed under a
profiler are never bench values."""
import argparse
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
import __graft_entry__ as g  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log-n", type=int, default=19)
    ap.add_argument("--cols", type=int, default=3)
    ap.add_argument("--sbox-d", type=int, default=5)
    ap.add_argument("--proves", type=int, default=1)
    a = ap.parse_args()
    pkg = g.load_package()
    ctx = pkg.Context(0)
    consts = bench.poseidon2_constants(0xB200, 8, 22)
    diag = np.stack([bench.ONE_MONT, bench.ONE_MONT, bench.TWO_MONT])
    ctx.check(ctx.lib.lsp_set_poseidon2(ctx.h, 3, a.sbox_d, 8, 22, pkg.ffi.as_u64p(consts), pkg.ffi.as_u64p(diag)), "set")
    n, c = 1 << a.log_n, a.cols
    pub = bench.random_fr_limbs(np.random.default_rng(7), 2)
    trace = ctx.permutation_trace(bench.synthetic_ab(0xB200, c, n), n, c, pub)
    cfgs = [pkg.AirPermutationConfig(range(c), range(c, 2 * c), 2 * c, 2 * c + 1)]
    for _ in range(a.proves):
        tm = {}
        pkg.prove(ctx, pkg.FriConfig(), cfgs, trace, pkg.from_mont_array(pub), timings=tm)
    print("stages_ms", {k: round(v, 2) for k, v in tm.items()}, "launches", ctx.kernel_launches())
    ctx.close()


if __name__ == "__main__":
    main()
