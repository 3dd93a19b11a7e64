"""Per-kernel times of the device verifier on the benchmark shape (3x3 columns, 2^LOG_N rows, blowup 8, 33 queries).
   python tools/verify_time.py [LOG_N]      (needs a GPU)"""
import sys
import time

import numpy as np

sys.path.insert(0, str(__import__("pathlib").Path(__file__).resolve().parent.parent))
import __graft_entry__ as g  # noqa: E402
import bench  # noqa: E402


def main():
    log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 19
    pkg = g.load_package()
    ctx = pkg.Context(0)
    consts = bench.poseidon2_constants(0xB200, 8, 22)
    diag = np.stack([bench.ONE_MONT, bench.ONE_MONT, bench.TWO_MONT])
    ctx.check(ctx.lib.lsp_set_poseidon2(ctx.h, 3, 5, 8, 22, pkg.ffi.as_u64p(consts), pkg.ffi.as_u64p(diag)), "lsp_set_poseidon2")
    n, c = 1 << log_n, 3
    fri = pkg.FriConfig()
    cfgs = [pkg.AirPermutationConfig(range(c), range(c, 2 * c), 2 * c, 2 * c + 1)]
    pub = bench.random_fr_limbs(np.random.default_rng(7), 2)
    trace = ctx.permutation_trace(bench.synthetic_ab(0xB200, c, n), n, c, pub)
    publics = pkg.from_mont_array(pub)
    proof = pkg.prove(ctx, fri, cfgs, trace, publics)
    pkg.verify(ctx, fri, cfgs, proof, publics)
    tm = {}
    t0 = time.perf_counter()
    for _ in range(10):
        pkg.verify(ctx, fri, cfgs, proof, publics, timing=tm)
    print(f"verify 2^{log_n} rows: wall {(time.perf_counter() - t0) * 100:.3f} ms per call, device {tm['device_ms']:.3f} ms")
    ctx.kernel_timing(True)
    pkg.verify(ctx, fri, cfgs, proof, publics)
    for r in ctx.kernel_timing_report():
        print(f"  {r['phase']:10s} {r['kernel']:28s} x{r['launches']:<3d} {r['ms']:.3f} ms")
    ctx.kernel_timing(False)
    bad = proof.words.copy()
    bad[-1] ^= np.uint64(1)
    print("tampered proof ->", pkg.VERIFY_REASONS[pkg.verify_code(ctx, fri, cfgs, bad, publics, log_n, proof.width)])


if __name__ == "__main__":
    main()
