"""The CUDA library against the frozen fixtures of tests/golden/golden_v1.json, through the C ABI."""
import pytest

from oracle import poseidon2 as OP
from tests.test_golden import GOLD, check_flat, instance, ix

pytestmark = pytest.mark.gpu


def test_field_vectors(gctx):
    a = [ix(v["a"]) for v in GOLD["field"]]
    b = [ix(v["b"]) for v in GOLD["field"]]
    for op in ("add", "sub", "mul"):
        assert gctx.fr_op(op, a, b) == [ix(v[op]) for v in GOLD["field"]]
    assert gctx.fr_op("inv", a) == [ix(v["inv_a"]) for v in GOLD["field"]]
    assert gctx.fr_op("halve", a) == [ix(v["halve_a"]) for v in GOLD["field"]]


@pytest.mark.parametrize("entry", GOLD["poseidon2"], ids=lambda e: f"d{e['sbox_d']}")
def test_poseidon2_vectors(pkg, entry):
    p = OP.Poseidon2Params.from_seed(entry["seed"], sbox_d=entry["sbox_d"], rounds_f=entry["rounds_f"], rounds_p=entry["rounds_p"])
    ctx = pkg.Context(0)
    ctx.set_poseidon2(p.sbox_d, p.rounds_f, p.rounds_p, p.flat_constants(), p.internal_diag_m1)
    assert ctx.permute([[ix(v) for v in c["in"]] for c in entry["cases"]]) == [[ix(v) for v in c["out"]] for c in entry["cases"]]
    ctx.close()


def test_sponge_lde_merkle_vectors(pkg, gctx):
    for c in GOLD["sponge"]:   # one width per call: hash_rows takes a matrix
        assert gctx.hash_rows([[ix(v) for v in c["row"]]] * 3) == [ix(c["digest"])] * 3
    d = gctx.upload([[ix(v) for v in r] for r in GOLD["lde"]["in"]])
    out = pkg.GpuDft(gctx).coset_lde_batch(d, GOLD["lde"]["added_bits"], ix(GOLD["lde"]["shift"]))
    assert out.rows() == [[ix(v) for v in r] for r in GOLD["lde"]["out_bitrev_storage"]]
    mm = pkg.GpuMmcs(gctx)
    root, tree = mm.commit([out])
    assert root == ix(GOLD["merkle"]["root"])
    for k, layer in enumerate(GOLD["merkle"]["layers"]):
        assert tree.layer(k) == [ix(v) for v in layer]
    assert mm.open_batch(5, tree)[1] == [ix(v) for v in GOLD["merkle"]["open_5"]["siblings"]]
    tree.free()
    out.free()
    d.free()


@pytest.mark.parametrize("case", GOLD["proofs"], ids=lambda c: f"2^{c['log_n']}x{c['cols']}")
def test_proof_vectors(pkg, gctx, case):
    cfgs, trace, publics = instance(case)
    g = [pkg.AirPermutationConfig(c.a_columns_ids, c.b_columns_ids, c.b_inverse_id, c.check_id) for c in cfgs]
    proof = pkg.prove(gctx, pkg.FriConfig(**case["fri"]), g, trace, publics)
    check_flat(case, proof.words)
    gd, idx = proof.to_dict()
    assert gd["commitments"]["trace"] == ix(case["trace_commit"]) and list(idx) == case["query_indices"]


def test_device_verifier_reproduces_the_frozen_rejection_reasons(pkg, gctx):
    """tests/golden/golden_verify_v1.json: one stored proof, 43 single-bit damages, the code each must be rejected with."""
    import numpy as np
    from tests.test_golden import verify_fixture
    case, cfgs, publics, words, vectors = verify_fixture()
    g = [pkg.AirPermutationConfig(x.a_columns_ids, x.b_columns_ids, x.b_inverse_id, x.check_id) for x in cfgs]
    fri = pkg.FriConfig(**case["fri"])
    w = sum(c.width() for c in g)
    assert pkg.verify_code(gctx, fri, g, words, publics, case["log_n"], w) == 0
    for word, bit, code in vectors:
        bad = words.copy()
        bad[word] ^= np.uint64(1 << bit)
        assert pkg.verify_code(gctx, fri, g, bad, publics, case["log_n"], w) == code, (word, bit)


def test_round2_fixture_on_the_device(pkg, p2params):
    """golden_round2_v1.json on the GPU: the proof hash under every non-default fork-only parameter, and a proof rebuilt from
    the committed serialised bytes is accepted by the device verifier."""
    from oracle import field as F
    from tests.test_golden import _fnv, _round2
    from tests.test_oracle_params import small_case
    g = _round2()
    ctx = pkg.Context(0)
    ctx.set_poseidon2(p2params.sbox_d, p2params.rounds_f, p2params.rounds_p, p2params.flat_constants(), p2params.internal_diag_m1)
    cfgs, trace, publics = small_case(seed=31, log_n=4, c=2)
    gc = [pkg.AirPermutationConfig(c.a_columns_ids, c.b_columns_ids, c.b_inverse_id, c.check_id) for c in cfgs]
    fri = pkg.FriConfig(**g["fri"])
    for name, v in g["params"].items():
        ctx.set_field_consts(int(v["generator"], 16), int(v["two_adic_root"], 16))
        ctx.set_transcript_flags(v["alpha_before_openings"], v["observe_opened_values"])
        assert _fnv(pkg.prove(ctx, fri, gc, trace, publics).words) == v["fnv1a64"], name
    ctx.set_field_consts(F.GENERATOR, F.TWO_ADIC_ROOT)
    ctx.set_transcript_flags()
    pkg.verify(ctx, fri, gc, pkg.Proof.deserialize(bytes.fromhex(g["serialized_hex"])), publics)
    ctx.close()
