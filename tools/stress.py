"""Determinism soak: the same proofs again and again (single GPU and local-sharded), and the lookup witness
under heavy key contention; any race in the grid barrier, the hash table or the scans shows up as a mismatch."""
import sys
import time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import __graft_entry__ as g
from oracle import field as F, trace as OT
from oracle.poseidon2 import Poseidon2Params

pkg = g.load_package()
p = Poseidon2Params.from_seed(0xB200, sbox_d=5)
ctx = pkg.Context(0)
ctx.set_poseidon2(p.sbox_d, p.rounds_f, p.rounds_p, p.flat_constants(), p.internal_diag_m1)
rng = F.SplitMix64(99)
alpha, delta = rng.next_fr(), rng.next_fr()
pub = pkg.to_mont_array([alpha, delta])
t0 = time.time()
bad = 0
for log_n, c in [(8, 2), (11, 3), (14, 1)]:
    n = 1 << log_n
    a, b = OT.synthetic_permutation_input(log_n, c, n)
    ab = pkg.to_mont_array([x for i in range(n) for x in [col[i] for col in a] + [col[i] for col in b]])
    dev = ctx.permutation_trace(ab, n, c, pub)
    cfgs = [pkg.AirPermutationConfig(range(c), range(c, 2 * c), 2 * c, 2 * c + 1)]
    fri = pkg.FriConfig(num_queries=20)
    ref = pkg.prove(ctx, fri, cfgs, dev, [alpha, delta]).words
    comm = pkg.Comm.local(ctx, 4)
    reps = 150 if log_n < 14 else 40
    for it in range(reps):
        w = pkg.prove(ctx, fri, cfgs, dev, [alpha, delta]).words
        bad += int(not np.array_equal(w, ref))
        if it % 5 == 0:
            w = pkg.prove_sharded(comm, fri, cfgs, dev, [alpha, delta]).words
            bad += int(not np.array_equal(w, ref))
    comm.close()
    print(f"2^{log_n} x {c}: {reps} repeats, mismatches so far {bad}", flush=True)
# lookup witness: 2^15 rows drawn from only 7 distinct table rows -> every atomic is contended
n = 1 << 15
a, b, af, bf = OT.synthetic_lookup_input(5, 2, 2, n, table_rows=7)
cols = list(a) + [col for t in b for col in t] + [af] + list(bf)
data = pkg.to_mont_array([x for i in range(n) for x in (col[i] for col in cols)])
ref = ctx.lookup_trace(data, n, 2, 2, 2, pub).download_array()
_, ocols = OT.lookup_columns(a, b, af, bf, alpha, delta)
assert np.array_equal(ref, pkg.to_mont_array([x for r in OT.row_major(ocols) for x in r])), "lookup witness differs from the oracle"
for it in range(60):
    bad += int(not np.array_equal(ctx.lookup_trace(data, n, 2, 2, 2, pub).download_array(), ref))
print(f"lookup witness 2^15 rows, 7 keys: 60 repeats, mismatches so far {bad}")
print(f"stress {'ok' if bad == 0 else 'FAILED'} in {time.time() - t0:.1f} s")
sys.exit(1 if bad else 0)
