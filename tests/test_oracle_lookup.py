"""LogUp lookup AIR in the oracle (no GPU): witness of `RawLookupTrace::get_trace`
(trace/src/lookup.rs:46-176), constraints of `eval_lookup` (air/src/lib.rs:57-114), quotient degree,
prove -> verify round trip, and the failure modes the reference asserts on."""
import pytest

from oracle import air as OA
from oracle import field as F
from oracle import stark as OS
from oracle import trace as OT


def test_lookup_witness_satisfies_air_and_sums_to_zero():
    rng = F.SplitMix64(1)
    alpha, delta = rng.next_fr(), rng.next_fr()
    lk = OT.synthetic_lookup_input(11, 2, 2, 32, disabled_every=5)
    cfg, cols = OT.lookup_columns(*lk, alpha, delta)
    assert cols[cfg.check_id][-1] == 0
    assert sum(cols[i][r] for i in cfg.occurrences_id for r in range(32)) == sum(lk[2])   # multiplicities count enabled A rows
    trace = OT.row_major(cols)
    assert len(trace[0]) == cfg.width()
    assert OA.check_constraints([cfg], trace, [alpha, delta])
    trace[3][cfg.occurrences_id[0]] = (trace[3][cfg.occurrences_id[0]] + 1) % F.R_MOD
    assert not OA.check_constraints([cfg], trace, [alpha, delta])


def test_lookup_of_missing_value_trips_the_assert():
    rng = F.SplitMix64(2)
    alpha, delta = rng.next_fr(), rng.next_fr()
    a, b, af, bf = OT.synthetic_lookup_input(12, 1, 1, 8)
    a[0][2] = 123456789          # not in the table
    with pytest.raises(AssertionError, match="check column should be 0"):
        OT.lookup_columns(a, b, af, bf, alpha, delta)
    af[2] = 0                    # ... unless the filter disables that row
    OT.lookup_columns(a, b, af, bf, alpha, delta)


def test_quotient_degree_and_constraint_count():
    lk = OA.AirLookupConfig.standard(2, 3, 2)
    pm = OA.AirPermutationConfig.standard(3, offset=lk.width())
    assert OA.log_quotient_degree([lk]) == 2 and OA.log_quotient_degree([lk, pm]) == 2
    assert OA.log_quotient_degree([OA.AirPermutationConfig.standard(3)]) == 1
    assert OA.num_constraints([lk, pm]) == (1 + 3 + 3) + 4


def test_prove_verify_lookup_and_permutation(p2params):
    rng = F.SplitMix64(3)
    alpha, delta = rng.next_fr(), rng.next_fr()
    cfgs, trace = OT.build_trace([OT.synthetic_permutation_input(4, 1, 16)], alpha, delta, [OT.synthetic_lookup_input(5, 2, 1, 16)])
    fri = OS.FriConfig(log_blowup=2, num_queries=4)
    proof = OS.prove(p2params, fri, cfgs, trace, [alpha, delta])
    assert len(proof["opened_values"]["quotient_chunks"]) == 4
    OS.verify(p2params, fri, cfgs, proof, [alpha, delta])
    proof["opened_values"]["trace_local"][cfgs[0].check_id] ^= 1
    with pytest.raises(OS.VerificationError):
        OS.verify(p2params, fri, cfgs, proof, [alpha, delta])


@pytest.mark.parametrize("log_n,lookups,perms,fri", [
    (3, [(1, 1, 0)], [], dict(log_blowup=2, log_final_poly_len=0, num_queries=5, proof_of_work_bits=0)),
    (5, [(2, 2, 6), (1, 1, 0)], [2], dict(log_blowup=3, log_final_poly_len=1, num_queries=9, proof_of_work_bits=2)),
])
def test_c_port_proves_and_verifies_lookup_airs(p2params, log_n, lookups, perms, fri):
    """The C port (oracle/c) pinned to the Python oracle on LineaAIRs with lookup configs: whole-proof equality."""
    import numpy as np
    from oracle import cport
    from tests.proofs import flat_from_dict
    cport.set_poseidon2(p2params)
    n = 1 << log_n
    rng = F.SplitMix64(log_n)
    alpha, delta = rng.next_fr(), rng.next_fr()
    lk = [OT.synthetic_lookup_input(30 + i, nc, nt, n, disabled_every=de) for i, (nc, nt, de) in enumerate(lookups)]
    pm = [OT.synthetic_permutation_input(40 + i, c, n) for i, c in enumerate(perms)]
    cfgs, trace = OT.build_trace(pm, alpha, delta, lk)
    ofri = OS.FriConfig(**fri)
    dbg = {}
    pp = OS.prove(p2params, ofri, cfgs, trace, [alpha, delta], dbg)
    words = cport.prove(ofri, cfgs, trace, [alpha, delta])
    assert np.array_equal(words, flat_from_dict(pp, dbg["query_indices"]))
    pub = np.array([F.to_mont_limbs(alpha), F.to_mont_limbs(delta)], dtype=np.uint64)
    w = OA.air_width(cfgs)
    assert cport.verify_limbs(ofri, log_n, w, cfgs, pub, words) == 0
    bad = words.copy()
    bad[4 * (2 + cfgs[0].check_id)] ^= 1
    assert cport.verify_limbs(ofri, log_n, w, cfgs, pub, bad) != 0
    # and the permutation-only path is unaffected by the registration above
    cfgs2, trace2 = OT.build_trace([OT.synthetic_permutation_input(5, 2, 8)], alpha, delta)
    assert cport.verify_limbs(ofri, 3, 6, cfgs2, pub, cport.prove(ofri, cfgs2, trace2, [alpha, delta])) == 0
