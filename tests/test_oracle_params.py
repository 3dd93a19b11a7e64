"""The parameters SURVEY.md 8(c) lists as unpinned -- the coset shift `Val::GENERATOR`, `two_adic_generator(47)`, and the
transcript order of `TwoAdicFriPcs::open` -- are run-time parameters of both restatements (Python oracle, C port), as
they are of the library (`lsp_set_field_consts`, `lsp_set_transcript_flags`).  For every non-default value: the two
restatements still produce the same proof, accept it, and it differs from the default-parameter proof."""
import numpy as np
import pytest

from oracle import air as OA
from oracle import cport
from oracle import field as F
from oracle import stark as OS
from oracle import trace as OT
from tests.proofs import flat_from_dict

ALT_GENERATOR = 5                                    # not a generator of Fr*, but outside every two-adic subgroup: a valid shift
ALT_ROOT = pow(F.TWO_ADIC_ROOT, 3, F.R_MOD)          # another primitive 2^47-th root of unity

CASES = {
    "generator": dict(consts=(ALT_GENERATOR, F.TWO_ADIC_ROOT)),
    "two_adic_root": dict(consts=(F.GENERATOR, ALT_ROOT)),
    "observe_then_alpha": dict(flags=(False, True)),   # upstream Plonky3 after early 2025
    "alpha_then_observe": dict(flags=(True, True)),
}


@pytest.fixture
def restore():
    yield
    F.set_field_consts()
    OS.set_transcript_flags()
    cport.set_field_consts(F.GENERATOR, F.TWO_ADIC_ROOT)
    cport.set_transcript_flags()


def apply_case(case):
    consts, flags = case.get("consts", (F.GENERATOR, F.TWO_ADIC_ROOT)), case.get("flags", (True, False))
    F.set_field_consts(*consts)
    OS.set_transcript_flags(*flags)
    cport.set_field_consts(*consts)
    cport.set_transcript_flags(*flags)
    return consts, flags


def small_case(seed=31, log_n=4, c=2):
    rng = F.SplitMix64(seed)
    alpha, delta = rng.next_fr(), rng.next_fr()
    cfgs, trace = OT.build_trace([OT.synthetic_permutation_input(seed + 1, c, 1 << log_n)], alpha, delta)
    return cfgs, trace, [alpha, delta]


@pytest.mark.parametrize("name", list(CASES))
def test_restatements_agree_for_every_non_default_parameter(p2params, restore, name):
    cport.set_poseidon2(p2params)
    fri = OS.FriConfig(log_blowup=2, log_final_poly_len=1, num_queries=5, proof_of_work_bits=2)
    cfgs, trace, publics = small_case()
    base = cport.prove(fri, cfgs, trace, publics)
    apply_case(CASES[name])
    dbg = {}
    proof = OS.prove(p2params, fri, cfgs, trace, publics, dbg)
    words = cport.prove(fri, cfgs, trace, publics)
    assert np.array_equal(words, flat_from_dict(proof, dbg["query_indices"]))
    assert not np.array_equal(words, base), "the parameter must reach the proof"
    OS.verify(p2params, fri, cfgs, proof, publics)
    pub = np.array([F.to_mont_limbs(x) for x in publics], dtype=np.uint64)
    assert cport.verify_limbs(fri, 4, OA.air_width(cfgs), cfgs, pub, words) == 0
    # ... and a verifier configured with the defaults rejects it
    F.set_field_consts()
    OS.set_transcript_flags()
    with pytest.raises(OS.VerificationError):
        OS.verify(p2params, fri, cfgs, proof, publics)


def test_alpha_after_openings_without_observing_is_the_default_transcript(p2params, restore):
    """Sampling after the openings changes nothing unless they are observed in between."""
    cport.set_poseidon2(p2params)
    fri = OS.FriConfig(log_blowup=1, log_final_poly_len=0, num_queries=3, proof_of_work_bits=0)
    cfgs, trace, publics = small_case(seed=77, log_n=3, c=1)
    base = cport.prove(fri, cfgs, trace, publics)
    cport.set_transcript_flags(False, False)
    assert np.array_equal(cport.prove(fri, cfgs, trace, publics), base)


def test_bad_field_constants_are_refused(restore):
    with pytest.raises(AssertionError):
        F.set_field_consts(F.GENERATOR, pow(F.TWO_ADIC_ROOT, 2, F.R_MOD))     # order 2^46 only
    with pytest.raises(AssertionError):
        F.set_field_consts(F.two_adic_generator(20), F.TWO_ADIC_ROOT)         # a shift inside H: g*H = H
    with pytest.raises(ValueError):
        cport.set_field_consts(F.GENERATOR, pow(F.TWO_ADIC_ROOT, 2, F.R_MOD))
