"""LogUp lookup AIR at benchmark size: device witness (`lsp_lookup_trace`) + prove (q = 4 chunks) + device verify for a
lookup of N_A columns into T tables at 2^LOG_N rows -- the AIR the reference's `main` proves at HEAD (bin/src/main.rs:37-43).
   python tools/lookup_bench.py [LOG_N] [N_COLS] [TABLES] [DISTINCT]     (needs a GPU)
DISTINCT limits the number of distinct looked-up rows (heavy key repetition, Linea-like); 0 = all rows distinct."""
import sys
import time

import numpy as np

sys.path.insert(0, str(__import__("pathlib").Path(__file__).resolve().parent.parent))
import __graft_entry__ as g  # noqa: E402
import bench  # noqa: E402


def main():
    log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 19
    nc = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    nt = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    distinct = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    n = 1 << log_n
    pkg = g.load_package()
    ctx = pkg.Context(0)
    consts = bench.poseidon2_constants(0xB200, 8, 22)
    diag = np.stack([bench.ONE_MONT, bench.ONE_MONT, bench.TWO_MONT])
    ctx.check(ctx.lib.lsp_set_poseidon2(ctx.h, 3, 5, 8, 22, pkg.ffi.as_u64p(consts), pkg.ffi.as_u64p(diag)), "lsp_set_poseidon2")
    rng = np.random.default_rng(5)
    pub = bench.random_fr_limbs(rng, 2)
    one = bench.ONE_MONT
    # table rows: `pool` distinct rows of nc elements; every table holds all of them (padded by repeats) in its own order;
    # the a side looks up random pool rows
    pool_n = distinct if distinct else n
    pool = bench.random_fr_limbs(rng, pool_n * nc).reshape(pool_n, nc, 4)
    stride = nc + nt * nc + 1 + nt
    rows = np.zeros((n, stride, 4), dtype=np.uint64)
    rows[:, :nc] = pool[rng.integers(0, pool_n, size=n)]
    for t in range(nt):
        order = np.concatenate([rng.permutation(pool_n), rng.integers(0, pool_n, size=n - pool_n)]) if pool_n < n else rng.permutation(n)
        rows[:, nc + t * nc: nc + (t + 1) * nc] = pool[order]
    rows[:, nc + nt * nc:] = one                                   # every filter enabled
    flat = np.ascontiguousarray(rows.reshape(n * stride, 4))
    ctx.sync()
    t0 = time.perf_counter()
    trace = ctx.lookup_trace(flat, n, nc, nt, nc, pub)
    ctx.sync()
    t_wit = (time.perf_counter() - t0) * 1e3
    warm = []
    for _ in range(3):
        t0 = time.perf_counter()
        trace2 = ctx.lookup_trace(flat, n, nc, nt, nc, pub)
        ctx.sync()
        warm.append((time.perf_counter() - t0) * 1e3)
        trace2.free()
    t_wit2 = min(warm)
    ctx.kernel_timing(True)
    trace2 = ctx.lookup_trace(flat, n, nc, nt, nc, pub)
    ctx.sync()
    report = ctx.kernel_timing_report()
    ctx.kernel_timing(False)
    trace2.free()
    base = 0
    cfg = pkg.AirLookupConfig(list(range(nc)), [[nc + t * nc + j for j in range(nc)] for t in range(nt)], nc + nt * nc,
                              [nc + nt * nc + 1 + t for t in range(nt)], nc + nt * nc + nt + 1,
                              [nc + nt * nc + nt + 2 + t for t in range(nt)], [nc + nt * nc + 2 * nt + 2 + t for t in range(nt)],
                              nc + nt * nc + 3 * nt + 2)
    assert cfg.width() == trace.width, (cfg.width(), trace.width)
    fri = pkg.FriConfig()
    publics = pkg.from_mont_array(pub)
    pkg.prove(ctx, fri, [cfg], trace, publics)
    tm = {}
    t0 = time.perf_counter()
    proof = pkg.prove(ctx, fri, [cfg], trace, publics, timings=tm)
    t_prove = (time.perf_counter() - t0) * 1e3
    vt = {}
    pkg.verify(ctx, fri, [cfg], proof, publics, timing=vt)
    print(f"lookup AIR: {nc} columns into {nt} tables, 2^{log_n} rows, {pool_n} distinct rows, trace width {trace.width}")
    print(f"  witness (upload of {flat.nbytes / 1e6:.0f} MB included): {t_wit:.1f} ms cold, {t_wit2:.1f} ms warm")
    print("  warm calls: " + ", ".join(f"{x:.1f}" for x in warm) + " ms; kernels: "
          + ", ".join(f"{r['kernel']} {r['ms']:.2f}" for r in sorted(report, key=lambda r: -r["ms"])[:8]))
    print(f"  prove: {t_prove:.1f} ms  " + ", ".join(f"{k} {v:.1f}" for k, v in tm.items()))
    print(f"  verify: accepted, {vt['device_ms']:.2f} ms")
    del base


if __name__ == "__main__":
    main()
