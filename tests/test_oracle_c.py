"""The C oracle (oracle/c) pinned to the Python oracle: field mul, Poseidon2, whole proofs,
and its verifier (accepts good proofs, names the failing check on tampered ones)."""
import numpy as np
import pytest

from oracle import air as OA
from oracle import cport
from oracle import field as F
from oracle import poseidon2 as OP
from oracle import stark as OS
from oracle import trace as OT

from tests.proofs import flat_from_dict


def test_c_field_and_permutation(p2params):
    rng = F.SplitMix64(1)
    a = [rng.next_fr() for _ in range(500)] + [0, 1, F.R_MOD - 1]
    b = [rng.next_fr() for _ in range(500)] + [F.R_MOD - 1, F.R_MOD - 1, F.R_MOD - 1]
    assert cport.fr_mul(a, b) == [x * y % F.R_MOD for x, y in zip(a, b)]
    for d in (3, 5, 7, 11, 17):
        p = OP.Poseidon2Params.from_seed(d, sbox_d=d)
        cport.set_poseidon2(p)
        st = [[rng.next_fr() for _ in range(3)] for _ in range(20)] + [[0, 0, 0]]
        assert cport.permute(st) == [OP.permute(p, s) for s in st]


@pytest.mark.parametrize("log_n,c,tables,fri", [
    (3, 3, 1, dict(log_blowup=3, log_final_poly_len=0, num_queries=33, proof_of_work_bits=0)),
    (5, 2, 2, dict(log_blowup=1, log_final_poly_len=0, num_queries=9, proof_of_work_bits=0)),
    (6, 6, 1, dict(log_blowup=2, log_final_poly_len=2, num_queries=7, proof_of_work_bits=3)),
])
def test_c_prove_equals_python_prove(p2params, log_n, c, tables, fri):
    cport.set_poseidon2(p2params)
    rng = F.SplitMix64(log_n)
    alpha, delta = rng.next_fr(), rng.next_fr()
    inputs = [OT.synthetic_permutation_input(5 + t, c, 1 << log_n) for t in range(tables)]
    cfgs, trace = OT.build_trace(inputs, alpha, delta)
    ofri = OS.FriConfig(**fri)
    dbg = {}
    pp = OS.prove(p2params, ofri, cfgs, trace, [alpha, delta], dbg)
    words = cport.prove(ofri, cfgs, trace, [alpha, delta])
    exp = flat_from_dict(pp, dbg["query_indices"])
    assert np.array_equal(words, exp)
    w = OA.air_width(cfgs)
    pub = np.array([F.to_mont_limbs(alpha), F.to_mont_limbs(delta)], dtype=np.uint64)
    assert cport.verify_limbs(ofri, log_n, w, cfgs, pub, words) == 0
    # tamper: opened value -> commit-phase/ input check fails; final poly -> transcript diverges
    bad = words.copy()
    bad[4 * 2] ^= 1
    assert cport.verify_limbs(ofri, log_n, w, cfgs, pub, bad) != 0
    bad = words.copy()
    bad[0] ^= 1
    assert cport.verify_limbs(ofri, log_n, w, cfgs, pub, bad) != 0


def test_c_gen_trace_satisfies_air():
    pub, tr, n, w = cport.gen_trace(0xB200, 3, 6)
    publics = [F.from_mont_limbs(r) for r in pub]
    flat = [F.from_mont_limbs(r) for r in tr]
    trace = [flat[i * w:(i + 1) * w] for i in range(n)]
    cfgs = [OA.AirPermutationConfig.standard(3)]
    assert OA.check_constraints(cfgs, trace, publics)
    assert trace[-1][-1] == 1
