"""`lsp_set_field_consts` / `lsp_set_transcript_flags` (the coset shift `Val::GENERATOR`,
`two_adic_generator(47)` and the transcript order of `TwoAdicFriPcs::open` are parameters, not constants).  One GPU case
per non-default value: the proof equals the CPU port's configured alike, the device verifier accepts it under the same
parameters and rejects it under the defaults; a sharded prove follows the parameters too."""
import numpy as np
import pytest

from oracle import air as OA
from oracle import cport
from oracle import field as F
from oracle import stark as OS
from tests.test_oracle_params import CASES, apply_case, restore, small_case  # noqa: F401  (restore is a fixture)

pytestmark = pytest.mark.gpu


def _gpu_cfgs(pkg, cfgs):
    return [pkg.AirPermutationConfig(c.a_columns_ids, c.b_columns_ids, c.b_inverse_id, c.check_id) for c in cfgs]


@pytest.fixture
def ctx(pkg, p2params):
    c = pkg.Context(0)
    c.set_poseidon2(p2params.sbox_d, p2params.rounds_f, p2params.rounds_p, p2params.flat_constants(), p2params.internal_diag_m1)
    yield c
    c.close()


@pytest.mark.parametrize("name", list(CASES))
def test_gpu_proof_follows_every_non_default_parameter(pkg, ctx, p2params, restore, name):
    cport.set_poseidon2(p2params)
    fri_kw = dict(log_blowup=2, log_final_poly_len=1, num_queries=5, proof_of_work_bits=2)
    cfgs, trace, publics = small_case(seed=41, log_n=6, c=2)
    g = _gpu_cfgs(pkg, cfgs)
    base = pkg.prove(ctx, pkg.FriConfig(**fri_kw), g, trace, publics)
    consts, flags = apply_case(CASES[name])
    ctx.set_field_consts(*consts)
    ctx.set_transcript_flags(*flags)
    mine = pkg.prove(ctx, pkg.FriConfig(**fri_kw), g, trace, publics)
    want = cport.prove(OS.FriConfig(**fri_kw), cfgs, trace, publics)
    assert np.array_equal(mine.words, want)
    assert not np.array_equal(mine.words, base.words)
    pkg.verify(ctx, pkg.FriConfig(**fri_kw), g, mine, publics)
    comm = pkg.Comm.local(ctx, 2)
    assert np.array_equal(pkg.prove_sharded(comm, pkg.FriConfig(**fri_kw), g, trace, publics).words, want)
    comm.close()
    # the default parameters again: the tables built from the other ones are gone, the old proof is back, the new one is rejected
    ctx.set_field_consts(F.GENERATOR, F.TWO_ADIC_ROOT)
    ctx.set_transcript_flags()
    assert np.array_equal(pkg.prove(ctx, pkg.FriConfig(**fri_kw), g, trace, publics).words, base.words)
    assert pkg.verify_code(ctx, pkg.FriConfig(**fri_kw), g, mine, publics) != 0
    pkg.verify(ctx, pkg.FriConfig(**fri_kw), g, base, publics)


def test_lde_follows_the_two_adic_root(pkg, ctx, restore):
    """`coset_lde_batch` alone, against the oracle's definition, under another primitive root."""
    from oracle import dft as OD
    alt = pow(F.TWO_ADIC_ROOT, 5, F.R_MOD)
    F.set_field_consts(F.GENERATOR, alt)
    ctx.set_field_consts(F.GENERATOR, alt)
    rng = F.SplitMix64(9)
    mat = [[rng.next_fr() for _ in range(3)] for _ in range(16)]
    got = pkg.GpuDft(ctx).coset_lde_batch(ctx.upload(mat), 2, 7)
    assert got.rows() == OD.coset_lde_batch(mat, 2, 7) == OD.coset_lde_batch_naive(mat, 2, 7)


def test_bad_field_constants_are_refused_and_leave_the_context_usable(pkg, ctx, p2params):
    cfgs, trace, publics = small_case(seed=43, log_n=4, c=1)
    g = _gpu_cfgs(pkg, cfgs)
    fri = pkg.FriConfig(log_blowup=1, num_queries=3)
    base = pkg.prove(ctx, fri, g, trace, publics)
    with pytest.raises(pkg.BackendError, match="primitive"):
        ctx.set_field_consts(F.GENERATOR, pow(F.TWO_ADIC_ROOT, 2, F.R_MOD))
    with pytest.raises(pkg.BackendError, match="generator"):
        ctx.set_field_consts(F.two_adic_generator(20), F.TWO_ADIC_ROOT)
    with pytest.raises(pkg.BackendError, match="generator"):
        ctx.set_field_consts(0, F.TWO_ADIC_ROOT)
    assert np.array_equal(pkg.prove(ctx, fri, g, trace, publics).words, base.words)
