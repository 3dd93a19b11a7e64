"""Size-independent properties at sizes the oracle cannot recompute: linearity of the LDE and of the FRI fold,
the LDE against its own coefficients at random points, Merkle openings against the root, and the witness
constraints on sampled rows -- all on limb arrays through the C ABI."""
import ctypes as C

import numpy as np
import pytest

from oracle import air as OA
from oracle import field as F
from oracle import merkle as OM

pytestmark = pytest.mark.gpu


def _rand_limbs(seed, n):
    """n uniform field elements as Montgomery limbs (any value below r is one)."""
    rng = np.random.default_rng(seed)
    out = rng.integers(0, 1 << 64, size=(n, 4), dtype=np.uint64)
    out[:, 3] &= np.uint64((1 << 60) - 1)          # < 2^252 < r
    return out


def _op(pkg, ctx, code, a, b):
    out = np.empty_like(a)
    ctx.check(ctx.lib.lsp_fr_op(ctx.h, code, pkg.ffi.as_u64p(a), pkg.ffi.as_u64p(b), pkg.ffi.as_u64p(out), len(a)), "lsp_fr_op")
    return out


def test_lde_is_linear_and_matches_its_coefficients(pkg, gctx):
    log_n, w, bits = 16, 4, 3
    n = 1 << log_n
    a, b = _rand_limbs(1, n * w), _rand_limbs(2, n * w)
    s = _op(pkg, gctx, 0, a, b)
    dft = pkg.GpuDft(gctx)
    la, ca = dft.coset_lde_batch(gctx.upload_array(a, n, w), bits, F.GENERATOR, want_coeffs=True)
    lb = dft.coset_lde_batch(gctx.upload_array(b, n, w), bits, F.GENERATOR)
    ls = dft.coset_lde_batch(gctx.upload_array(s, n, w), bits, F.GENERATOR)
    A, B, S = la.download_array(), lb.download_array(), ls.download_array()
    assert np.array_equal(_op(pkg, gctx, 0, A, B), S)                       # LDE(a + b) = LDE(a) + LDE(b)
    # storage row p of column c is p_c(g * w_L^bitrev(p)): Horner over the returned coefficients at a few rows
    coef = pkg.from_mont_array(ca.download_array().reshape(n, w, 4)[:, 1, :].copy())
    log_l = log_n + bits
    w_l = F.two_adic_generator(log_l)
    for p in (0, 1, 12345, (1 << log_l) - 1):
        x = F.GENERATOR * pow(w_l, F.reverse_bits_len(p, log_l), F.R_MOD) % F.R_MOD
        acc = 0
        for ck in reversed(coef):
            acc = (acc * x + ck) % F.R_MOD
        assert pkg.from_mont_array(A.reshape(-1, w, 4)[p:p + 1, 1, :].copy())[0] == acc


def test_fri_fold_is_linear(pkg, gctx):
    n = 1 << 18
    u, v = _rand_limbs(3, n), _rand_limbs(4, n)
    beta = F.SplitMix64(5).next_fr()
    fold = lambda x: pkg.fri_fold(gctx, gctx.upload_array(x, n, 1), beta).download_array()
    assert np.array_equal(_op(pkg, gctx, 0, fold(u), fold(v)), fold(_op(pkg, gctx, 0, u, v)))


def test_merkle_openings_verify_at_2p18_leaves(pkg, gctx, p2params):
    h, w = 1 << 18, 3
    m = gctx.upload_array(_rand_limbs(6, h * w), h, w)
    mm = pkg.GpuMmcs(gctx)
    root, tree = mm.commit([m])
    for idx in (0, 1, h // 3, h - 1):
        rows, proof = mm.open_batch(idx, tree)
        assert len(proof) == 18
        assert OM.verify_batch(p2params, root, h, idx, rows, proof)
        bad = [list(r) for r in rows]
        bad[0][0] = (bad[0][0] + 1) % F.R_MOD
        assert not OM.verify_batch(p2params, root, h, idx, bad, proof)
    tree.free()


def test_witness_satisfies_constraints_on_sampled_rows_at_2p19(pkg, gctx):
    """The 2^19-row benchmark witness: every sampled row pair satisfies `eval_permutation`, the last row closes the product."""
    import bench
    log_n, c = 19, 3
    n, w = 1 << log_n, 2 * c + 2
    pub = bench.random_fr_limbs(np.random.default_rng(7), 2)
    dev = gctx.permutation_trace(bench.synthetic_ab(0xB200, c, n), n, c, pub)
    publics = pkg.from_mont_array(pub)
    cfgs = [OA.AirPermutationConfig.standard(c)]
    rng = np.random.default_rng(8)
    for i in [0, n - 2] + [int(x) for x in rng.integers(1, n - 2, size=12)]:
        two = dev.rows(i, 2)
        cs = OA.eval_air(cfgs, [OA.Fe(x) for x in two[0]], [OA.Fe(x) for x in two[1]], [OA.Fe(x) for x in publics], OA.Fe(0), OA.Fe(1),
                         OA.Fe(1 if i == 0 else 0), OA.Fe(0), OA.Fe(1))
        assert all(x.v == 0 for x in cs), i
    assert dev.rows(n - 1, 1)[0][-1] == 1


def test_random_shapes_against_the_c_port(pkg, gctx, p2params):
    """Differential sweep: 40 seeded random shapes (height 2..2^12, 1..6 columns, blowup 2..8, final-poly length 1..4,
    0..5 proof-of-work bits, 1..40 queries): the GPU proof must equal the C port's word for word and verify."""
    from oracle import cport
    from oracle import stark as OS
    cport.set_poseidon2(p2params)
    rng = np.random.default_rng(2026)
    done = refused = 0
    for trial in range(60):
        log_n = int(rng.integers(1, 13))
        c = int(rng.integers(1, 7))
        log_blowup = int(rng.integers(1, 4))
        log_final = int(rng.integers(0, min(3, log_n + 1)))
        fri_kw = dict(log_blowup=log_blowup, log_final_poly_len=log_final, num_queries=int(rng.integers(1, 41)),
                      proof_of_work_bits=int(rng.integers(0, 6)))
        if log_blowup + log_final > 10:
            continue
        pub, tr, n, w = cport.gen_trace(int(rng.integers(1, 1 << 30)), c, log_n)
        cfgs = [OA.AirPermutationConfig.standard(c)]
        g = [pkg.AirPermutationConfig(x.a_columns_ids, x.b_columns_ids, x.b_inverse_id, x.check_id) for x in cfgs]
        ofri = OS.FriConfig(**fri_kw)
        if log_final == log_n:
            # Zero commit-phase rounds (final polynomial as long as the trace): the reduced opening never enters the fold
            # chain, so the low-degree test is vacuous -- the pinned verifier rejects every honest proof
            # (FinalPolyMismatch) and would accept a forged all-zero final polynomial.  The library refuses the
            # configuration on both sides instead of producing or checking such a proof.
            with pytest.raises(pkg.BackendError, match="no commit-phase round"):
                pkg.prove(gctx, pkg.FriConfig(**fri_kw), g, (tr, n, w), pkg.from_mont_array(pub))
            refused += 1
            continue
        cwords = cport.prove_limbs(ofri, tr, n, w, cfgs, pub)
        gproof = pkg.prove(gctx, pkg.FriConfig(**fri_kw), g, (tr, n, w), pkg.from_mont_array(pub))
        assert np.array_equal(cwords, gproof.words), (log_n, c, fri_kw)
        assert cport.verify_limbs(ofri, log_n, w, cfgs, pub, gproof.words) == 0, (log_n, c, fri_kw)
        assert pkg.verify_code(gctx, pkg.FriConfig(**fri_kw), g, gproof, pkg.from_mont_array(pub)) == 0, (log_n, c, fri_kw)
        done += 1
        if done == 40:
            break
    assert done >= 30 and refused >= 1


@pytest.mark.parametrize("bits", [16, 20])
def test_grinding_at_real_bit_counts_finds_the_cpu_ports_witness(pkg, gctx, p2params, bits):
    """`proof_of_work_bits` at sizes where grinding is real work (the reference's commented setting is 29,
    bin/src/main.rs:62; timed by tools/grind_bench.py): the device grinder -- one permutation per trial on top of the
    shared sponge prefix -- must find the SMALLEST witness, i.e. the one the CPU port's ordered search finds, for an odd
    and an even number of buffered observations (final polynomial of 2 resp. 1 coefficients ahead of the witness)."""
    from oracle import cport
    from oracle import stark as OS
    cport.set_poseidon2(p2params)
    cport.set_threads(0)
    for log_blowup, log_final in ((1, 0), (1, 1), (2, 0)):
        fri_kw = dict(log_blowup=log_blowup, log_final_poly_len=log_final, num_queries=3, proof_of_work_bits=bits)
        pub, tr, n, w = cport.gen_trace(77 + bits + log_final, 1, 3)
        cfgs = [OA.AirPermutationConfig.standard(1)]
        g = [pkg.AirPermutationConfig(x.a_columns_ids, x.b_columns_ids, x.b_inverse_id, x.check_id) for x in cfgs]
        gproof = pkg.prove(gctx, pkg.FriConfig(**fri_kw), g, (tr, n, w), pkg.from_mont_array(pub))
        assert np.array_equal(cport.prove_limbs(OS.FriConfig(**fri_kw), tr, n, w, cfgs, pub), gproof.words), fri_kw
        assert pkg.verify_code(gctx, pkg.FriConfig(**fri_kw), g, gproof, pkg.from_mont_array(pub)) == 0
        d, _ = gproof.to_dict()
        assert d["opening_proof"]["pow_witness"] > 0
