#!/usr/bin/env python
"""Proof-of-work grinding at the reference's commented setting (`proof_of_work_bits: 0, //29`, bin/src/main.rs:62; 19.3 s
of its bench.log:66): full proves of the bench workload with 0 / 20 / 29 bits, the grind+query stage time, the witness found
and the trial rate of k_ch_grind_chunk (one Poseidon2 permutation per trial) against the compress-layer rate.
    python tools/grind_bench.py [log_n] [bits ...]"""
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
import __graft_entry__ as g  # noqa: E402


def main():
    log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 19
    bits_list = [int(x) for x in sys.argv[2:]] or [0, 16, 20, 24, 29]
    pkg = g.load_package()
    ctx = pkg.Context(0)
    args = type("A", (), dict(log_n=log_n, cols=3, log_blowup=3, sbox_d=5))()
    ab, pub, consts, diag = bench.workload_inputs(args)
    ctx.check(ctx.lib.lsp_set_poseidon2(ctx.h, 3, 5, 8, 22, pkg.ffi.as_u64p(consts), pkg.ffi.as_u64p(diag)), "set")
    n, c = 1 << log_n, 3
    trace = ctx.permutation_trace(ab, n, c, pub)
    cfgs = [pkg.AirPermutationConfig(range(c), range(c, 2 * c), 2 * c, 2 * c + 1)]
    publics = pkg.from_mont_array(pub)
    rows = []
    for bits in bits_list:
        fri = pkg.FriConfig(log_blowup=3, log_final_poly_len=0, num_queries=33, proof_of_work_bits=bits)
        pkg.prove(ctx, fri, cfgs, trace, publics)                       # warm
        tm = {}
        t0 = time.perf_counter()
        proof = pkg.prove(ctx, fri, cfgs, trace, publics, timings=tm)
        wall = time.perf_counter() - t0
        pkg.verify(ctx, fri, cfgs, proof, publics)
        d, _ = proof.to_dict()
        wit = int(d["opening_proof"]["pow_witness"])
        chunk = 1 << min(25, max(17, bits + 1))
        trials = ((wit // chunk) + 1) * chunk if bits else 0
        grind_ms = tm["grind_query"]
        rows.append(dict(bits=bits, prove_s=round(wall, 4), grind_query_ms=round(grind_ms, 3), witness=wit, trials_tested=trials,
                         trials_per_s=round(trials / (grind_ms * 1e-3)) if bits else None, verified=True))
        print(json.dumps(rows[-1]), flush=True)
    base = rows[0]["grind_query_ms"] if rows and rows[0]["bits"] == 0 else 0.0
    for r in rows:
        if r["bits"]:
            print(f"# {r['bits']} bits: grind alone ~{r['grind_query_ms'] - base:.1f} ms for {r['trials_tested']} trials "
                  f"= {r['trials_tested'] / max(1e-9, (r['grind_query_ms'] - base) * 1e-3) / 1e6:.0f} M trials/s "
                  f"(compress layers run at ~535 M permutations/s; the reference's CPU grind span is 19.3 s, bench.log:66)")
    ctx.close()


if __name__ == "__main__":
    main()
