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


# Python oracle's error names <-> the C port's codes (and LSP_VERIFY_* of include/lsp_b200.h)
_REASON_CODES = {"InvalidProofShape": {1}, "InputError(MerkleRootMismatch)": {2, 3}, "CommitPhaseMmcsError": {4},
                 "FinalPolyMismatch": {5}, "InvalidPowWitness": {6}, "OodEvaluationMismatch": {7}}


def test_c_verifier_names_the_same_failing_check_as_the_python_verifier(p2params):
    """The two restatements of `verify` must agree not only on accept/reject but on WHICH check fails, for every kind
    of damage a structured proof can carry (the device verifier is then compared with the C port code for code)."""
    import copy
    cport.set_poseidon2(p2params)
    log_n, c = 5, 2
    fri = dict(log_blowup=2, log_final_poly_len=1, num_queries=6, proof_of_work_bits=3)
    rng = F.SplitMix64(99)
    alpha, delta = rng.next_fr(), rng.next_fr()
    cfgs, trace = OT.build_trace([OT.synthetic_permutation_input(9, c, 1 << log_n)], alpha, delta)
    ofri = OS.FriConfig(**fri)
    dbg = {}
    good = OS.prove(p2params, ofri, cfgs, trace, [alpha, delta], dbg)
    idx = dbg["query_indices"]
    w = OA.air_width(cfgs)
    pub = np.array([F.to_mont_limbs(alpha), F.to_mont_limbs(delta)], dtype=np.uint64)

    def both(proof, use_cfgs=cfgs):
        try:
            OS.verify(p2params, ofri, use_cfgs, proof, [alpha, delta])
            name = None
        except OS.VerificationError as e:
            name = str(e)
        code = cport.verify_limbs(ofri, log_n, w, use_cfgs, pub, flat_from_dict(proof, idx))
        return name, code

    assert both(good) == (None, 0)
    bump = lambda v: (v + 1) % F.R_MOD
    damage = {}
    p = copy.deepcopy(good); q = p["opening_proof"]["query_proofs"][2]
    q["input_proof"][0]["opened_values"][0][1] = bump(q["input_proof"][0]["opened_values"][0][1]); damage["trace row"] = p
    p = copy.deepcopy(good); q = p["opening_proof"]["query_proofs"][0]
    q["input_proof"][0]["opening_proof"][-1] = bump(q["input_proof"][0]["opening_proof"][-1]); damage["trace path"] = p
    p = copy.deepcopy(good); q = p["opening_proof"]["query_proofs"][5]
    q["input_proof"][1]["opened_values"][1][0] = bump(q["input_proof"][1]["opened_values"][1][0]); damage["quotient row"] = p
    p = copy.deepcopy(good); q = p["opening_proof"]["query_proofs"][1]
    q["commit_phase_openings"][0]["sibling_value"] = bump(q["commit_phase_openings"][0]["sibling_value"]); damage["fri sibling"] = p
    p = copy.deepcopy(good); q = p["opening_proof"]["query_proofs"][3]
    q["commit_phase_openings"][-1]["opening_proof"][0] = bump(q["commit_phase_openings"][-1]["opening_proof"][0]); damage["fri path"] = p
    p = copy.deepcopy(good); p["opened_values"]["trace_next"][0] = bump(p["opened_values"]["trace_next"][0]); damage["opened value"] = p
    p = copy.deepcopy(good); p["opening_proof"]["pow_witness"] = bump(p["opening_proof"]["pow_witness"]); damage["pow witness"] = p
    seen = set()
    for what, proof in damage.items():
        name, code = both(proof)
        assert name is not None and code in _REASON_CODES[name], (what, name, code)
        seen.add(name)
    c0 = cfgs[0]
    swapped = [OA.AirPermutationConfig(list(reversed(c0.a_columns_ids)), c0.b_columns_ids, c0.b_inverse_id, c0.check_id)]
    name, code = both(good, swapped)
    assert (name, code) == ("OodEvaluationMismatch", 7)
    assert {"InputError(MerkleRootMismatch)", "CommitPhaseMmcsError"} <= seen


def test_zero_round_fri_is_rejected_at_the_final_polynomial_by_both_verifiers(p2params):
    """log_final_poly_len == log2(height): the provers agree, and both verifiers -- like the pinned Plonky3 one in a
    release build -- never add the reduced opening (that happens inside a folding round) and report FinalPolyMismatch."""
    cport.set_poseidon2(p2params)
    for log_n, log_blowup in ((1, 1), (3, 2)):
        fri = OS.FriConfig(log_blowup=log_blowup, log_final_poly_len=log_n, num_queries=4, proof_of_work_bits=0)
        rng = F.SplitMix64(50 + log_n)
        alpha, delta = rng.next_fr(), rng.next_fr()
        cfgs, trace = OT.build_trace([OT.synthetic_permutation_input(3, 2, 1 << log_n)], alpha, delta)
        dbg = {}
        proof = OS.prove(p2params, fri, cfgs, trace, [alpha, delta], dbg)
        words = cport.prove(fri, cfgs, trace, [alpha, delta])
        assert np.array_equal(words, flat_from_dict(proof, dbg["query_indices"]))
        assert proof["opening_proof"]["commit_phase_commits"] == []
        with pytest.raises(OS.VerificationError, match="FinalPolyMismatch"):
            OS.verify(p2params, fri, cfgs, proof, [alpha, delta])
        pub = np.array([F.to_mont_limbs(alpha), F.to_mont_limbs(delta)], dtype=np.uint64)
        assert cport.verify_limbs(fri, log_n, OA.air_width(cfgs), cfgs, pub, words) == 5
