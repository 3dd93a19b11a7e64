"""GPU parity for the STARK stages, each through the C ABI, bit-exact against the oracle:
coset LDE (K2/K3), quotient (K6), FRI fold (K9), and the full prove (transcript,
openings, FRI, queries) -- plus acceptance of the GPU proof by the oracle's verifier."""
import copy

import pytest

from oracle import air as OA
from oracle import dft as OD
from oracle import field as F
from oracle import poseidon2 as OP
from oracle import stark as OS
from oracle import trace as OT

pytestmark = pytest.mark.gpu


def _rand_mat(n, w, seed):
    rng = F.SplitMix64(seed)
    return [[rng.next_fr() for _ in range(w)] for _ in range(n)]


@pytest.mark.parametrize("log_n,w,bits", [(0, 3, 3), (1, 1, 1), (2, 2, 2), (3, 8, 3), (5, 3, 1), (6, 14, 3), (9, 2, 2),
                                          (10, 1, 3), (11, 3, 1), (12, 1, 2)])
def test_coset_lde_batch(pkg, gctx, log_n, w, bits):
    mat = _rand_mat(1 << log_n, w, 1000 + log_n * 7 + w)
    for shift in (F.GENERATOR, pow(F.two_adic_generator(log_n + 1), (1 << (log_n + 1)) - 1, F.R_MOD)):
        exp = OD.coset_lde_batch(mat, bits, shift)
        d = gctx.upload(mat)
        out, co = pkg.GpuDft(gctx).coset_lde_batch(d, bits, shift, want_coeffs=True)
        assert out.height == (1 << (log_n + bits)) and out.width == w
        assert out.rows() == exp
        assert co.rows() == OD.rows_of([OD.idft(c) for c in OD.columns_of(mat)])
        for m in (d, out, co):
            m.free()


def test_coset_lde_matches_definition(pkg, gctx):
    mat = _rand_mat(8, 2, 5)
    d = gctx.upload(mat)
    out = pkg.GpuDft(gctx).coset_lde_batch(d, 2, F.GENERATOR)
    assert out.rows() == OD.coset_lde_batch_naive(mat, 2, F.GENERATOR)


def test_coset_lde_rejects_ragged_height(pkg, gctx):
    d = gctx.upload(_rand_mat(6, 2, 1))
    with pytest.raises(pkg.BackendError):
        pkg.GpuDft(gctx).coset_lde_batch(d, 1, F.GENERATOR)


def _perm_instance(log_n, c, seed, n_tables=1):
    rng = F.SplitMix64(seed)
    alpha, delta = rng.next_fr(), rng.next_fr()
    inputs = [OT.synthetic_permutation_input(seed + 17 * t, c, 1 << log_n) for t in range(n_tables)]
    cfgs, trace = OT.build_trace(inputs, alpha, delta)
    assert OA.check_constraints(cfgs, trace, [alpha, delta])
    return cfgs, trace, [alpha, delta]


def _gpu_cfgs(pkg, cfgs):
    return [pkg.AirPermutationConfig(c.a_columns_ids, c.b_columns_ids, c.b_inverse_id, c.check_id) for c in cfgs]


@pytest.mark.parametrize("log_n,c,tables", [(1, 1, 1), (3, 3, 1), (5, 2, 1), (6, 3, 2), (8, 6, 1)])
def test_quotient_values(pkg, gctx, log_n, c, tables):
    cfgs, trace, publics = _perm_instance(log_n, c, 40 + log_n, tables)
    log_q, bits = 1, 3
    lde = OD.coset_lde_batch(trace, bits, F.GENERATOR)
    td, qd = OS.Domain(log_n, 1), OS.Domain(log_n + log_q, F.GENERATOR)
    alpha = F.SplitMix64(99).next_fr()
    toq = OD.bit_reverse_rows(lde[:qd.size()])
    qv = OS.quotient_values(cfgs, publics, td, qd, toq, alpha)
    d = gctx.upload(lde)
    got = pkg.quotient_permutation(gctx, d, log_n, log_q, _gpu_cfgs(pkg, cfgs), publics, alpha).rows()
    exp = [[qv[k * 2 + ch] for ch in range(2)] for k in range(1 << log_n)]
    assert got == exp


@pytest.mark.parametrize("log_len", [1, 2, 3, 7, 12])
def test_fri_fold(pkg, gctx, log_len):
    rng = F.SplitMix64(log_len)
    v = [rng.next_fr() for _ in range(1 << log_len)]
    beta = rng.next_fr()
    exp = OS.fold_matrix(beta, [[v[2 * j], v[2 * j + 1]] for j in range(len(v) // 2)])
    got = pkg.fri_fold(gctx, gctx.upload([[x] for x in v]), beta).rows()
    assert [r[0] for r in got] == exp
    # fold_row (the verifier's definition) agrees at a few indices
    for j in {0, len(exp) - 1, len(exp) // 3}:
        assert OS.fold_row(j, log_len - 1, beta, v[2 * j], v[2 * j + 1]) == exp[j]


@pytest.mark.parametrize("log_n,c,tables,fri", [
    (3, 3, 1, dict(log_blowup=3, log_final_poly_len=0, num_queries=33, proof_of_work_bits=0)),
    (5, 3, 1, dict(log_blowup=3, log_final_poly_len=0, num_queries=33, proof_of_work_bits=0)),
    (6, 2, 2, dict(log_blowup=1, log_final_poly_len=0, num_queries=10, proof_of_work_bits=0)),
    (6, 6, 1, dict(log_blowup=2, log_final_poly_len=2, num_queries=7, proof_of_work_bits=0)),
    (4, 1, 1, dict(log_blowup=3, log_final_poly_len=0, num_queries=5, proof_of_work_bits=6)),
    (8, 3, 1, dict(log_blowup=3, log_final_poly_len=0, num_queries=33, proof_of_work_bits=0)),
])
def test_prove_bit_exact_and_verifies(pkg, gctx, p2params, log_n, c, tables, fri):
    cfgs, trace, publics = _perm_instance(log_n, c, 7 + log_n + c, tables)
    ofri = OS.FriConfig(**fri)
    dbg = {}
    oproof = OS.prove(p2params, ofri, cfgs, trace, publics, dbg)
    OS.verify(p2params, ofri, cfgs, oproof, publics)
    gproof = pkg.prove(gctx, pkg.FriConfig(**fri), _gpu_cfgs(pkg, cfgs), trace, publics)
    gd, indices = gproof.to_dict()
    assert gd["commitments"] == oproof["commitments"]
    assert gd["opened_values"] == oproof["opened_values"]
    assert gd["opening_proof"]["commit_phase_commits"] == oproof["opening_proof"]["commit_phase_commits"]
    assert gd["opening_proof"]["final_poly"] == oproof["opening_proof"]["final_poly"]
    assert gd["opening_proof"]["pow_witness"] == oproof["opening_proof"]["pow_witness"]
    assert indices == dbg["query_indices"]
    assert gd == oproof
    # the (restated) unchanged verifier accepts the GPU proof and rejects a tampered one
    OS.verify(p2params, ofri, cfgs, gd, publics)
    bad = copy.deepcopy(gd)
    bad["opened_values"]["trace_next"][0] = (bad["opened_values"]["trace_next"][0] + 1) % F.R_MOD
    with pytest.raises(OS.VerificationError):
        OS.verify(p2params, ofri, cfgs, bad, publics)


def test_prove_rejects_bad_shapes(pkg, gctx):
    cfgs, trace, publics = _perm_instance(3, 2, 3)
    g = _gpu_cfgs(pkg, cfgs)
    with pytest.raises(pkg.BackendError):
        pkg.prove(gctx, pkg.FriConfig(), g, trace[:6], publics)             # not a power of two
    with pytest.raises(pkg.BackendError):
        pkg.prove(gctx, pkg.FriConfig(log_blowup=0), g, trace, publics)     # quotient degree > blowup
    with pytest.raises(pkg.BackendError):
        pkg.prove(gctx, pkg.FriConfig(), g, [r[:-1] for r in trace], publics)  # AIR width != trace width


@pytest.mark.parametrize("log_n,c", [(0, 1), (3, 3), (7, 2), (11, 3), (13, 1)])
def test_permutation_witness_on_device(pkg, gctx, log_n, c):
    """lsp_permutation_trace == RawPermutationTrace::get_trace + RawTrace::get_trace."""
    import numpy as np
    n = 1 << log_n
    rng = F.SplitMix64(70 + log_n)
    alpha, delta = rng.next_fr(), rng.next_fr()
    a, b = OT.synthetic_permutation_input(9 + log_n, c, n)
    _, cols = OT.permutation_columns(a, b, alpha, delta)
    exp = OT.row_major(cols)
    ab = pkg.to_mont_array([x for i in range(n) for x in [col[i] for col in a] + [col[i] for col in b]])
    got = gctx.permutation_trace(ab, n, c, pkg.to_mont_array([alpha, delta]))
    assert got.rows() == exp
    # a non-permutation trips the reference's last-row assertion
    if n > 1:
        bad = ab.copy()
        bad[0] = pkg.to_mont_array([12345])[0]
        with pytest.raises(pkg.BackendError, match="check column should be 1"):
            gctx.permutation_trace(bad, n, c, pkg.to_mont_array([alpha, delta]))


def test_full_size_prove_verifies_and_matches_cpu_port(pkg):
    """BASELINE configs[1]: 3x3 columns, 2^19 rows.  The GPU proof must be accepted by the
    (restated) verifier and be bit-identical to the multi-threaded CPU port's proof."""
    import os
    import numpy as np
    from oracle import cport
    from oracle.poseidon2 import Poseidon2Params
    log_n = int(os.environ.get("LSP_FULL_LOG_N", "19"))
    c = 3
    p = Poseidon2Params.from_seed(0xB200, sbox_d=5)
    cport.set_poseidon2(p)
    ctx = pkg.Context(0)
    ctx.set_poseidon2(p.sbox_d, p.rounds_f, p.rounds_p, p.flat_constants(), p.internal_diag_m1)
    pub, tr, n, w = cport.gen_trace(0xB200, c, log_n)
    fri = OS.FriConfig()
    cfgs = [OA.AirPermutationConfig.standard(c)]
    gproof = pkg.prove(ctx, pkg.FriConfig(), _gpu_cfgs(pkg, cfgs), (tr, n, w), pkg.from_mont_array(pub))
    assert cport.verify_limbs(fri, log_n, w, cfgs, pub, gproof.words) == 0
    pkg.verify(ctx, pkg.FriConfig(), _gpu_cfgs(pkg, cfgs), gproof, pkg.from_mont_array(pub))      # the device verifier accepts it too
    bad = gproof.words.copy()
    bad[4 * 5] ^= 1
    why = cport.verify_limbs(fri, log_n, w, cfgs, pub, bad)
    assert why != 0
    assert pkg.verify_code(ctx, pkg.FriConfig(), _gpu_cfgs(pkg, cfgs), bad, pkg.from_mont_array(pub), log_n, w) == why
    cwords = cport.prove_limbs(fri, tr, n, w, cfgs, pub)
    assert np.array_equal(cwords, gproof.words)
    # the device-side witness generator reproduces the CPU port's trace from its a/b columns
    ab = np.ascontiguousarray(tr.reshape(n, w, 4)[:, :2 * c, :].reshape(n * 2 * c, 4))
    assert np.array_equal(ctx.permutation_trace(ab, n, c, pub).download_array(), tr)
    ctx.close()


@pytest.mark.parametrize("name,log_n,c,log_blowup,port_seconds", [
    ("cfg3: 3x3 columns, 2^22 rows, blowup 8", 22, 3, 3, 210),
    ("cfg4a: 3x32 columns, 2^20 rows, blowup 2", 20, 32, 1, 55),
    ("cfg4b: 3x32 columns, 2^20 rows, blowup 4", 20, 32, 2, 110),
])
def test_baseline_configs_bit_exact(pkg, name, log_n, c, log_blowup, port_seconds):
    """BASELINE.json configs[2] and configs[3] as parity cases: the trace comes from the CPU port's generator
    (trace/src/permutation.rs semantics), the device witness generator must reproduce it, the GPU proof must be accepted
    by both verifiers (a flipped bit rejected for the same reason) -- and it must BE the CPU port's proof, word for word.
    The port needs 1-4 minutes of host time at these sizes: cfg4a always runs in full; the word-for-word leg of cfg3 and
    cfg4b runs under LSP_FULL_PARITY=1 (recorded per round in profiles/), their acceptance legs always."""
    import os
    import numpy as np
    from oracle import cport
    from oracle.poseidon2 import Poseidon2Params
    if os.environ.get("LSP_SKIP_BIG"):
        pytest.skip("LSP_SKIP_BIG set")
    p = Poseidon2Params.from_seed(0xB200, sbox_d=5)
    cport.set_poseidon2(p)
    cport.set_threads(0)
    ctx = pkg.Context(0)
    ctx.set_poseidon2(p.sbox_d, p.rounds_f, p.rounds_p, p.flat_constants(), p.internal_diag_m1)
    pub, tr, n, w = cport.gen_trace(0xC0FFEE + log_n, c, log_n)
    assert w == 2 * c + 2
    ab = np.ascontiguousarray(tr.reshape(n, w, 4)[:, :2 * c, :].reshape(n * 2 * c, 4))
    dev = ctx.permutation_trace(ab, n, c, pub)
    assert np.array_equal(dev.download_array(), tr)
    fri_kw = dict(log_blowup=log_blowup, log_final_poly_len=0, num_queries=33, proof_of_work_bits=0)
    cfgs = [OA.AirPermutationConfig.standard(c)]
    gproof = pkg.prove(ctx, pkg.FriConfig(**fri_kw), _gpu_cfgs(pkg, cfgs), dev, pkg.from_mont_array(pub))
    ofri = OS.FriConfig(**fri_kw)
    assert cport.verify_limbs(ofri, log_n, w, cfgs, pub, gproof.words) == 0
    pkg.verify(ctx, pkg.FriConfig(**fri_kw), _gpu_cfgs(pkg, cfgs), gproof, pkg.from_mont_array(pub))
    bad = gproof.words.copy()
    bad[4 * 3 + 1] ^= 1 << 7
    why = cport.verify_limbs(ofri, log_n, w, cfgs, pub, bad)
    assert why != 0
    assert pkg.verify_code(ctx, pkg.FriConfig(**fri_kw), _gpu_cfgs(pkg, cfgs), bad, pkg.from_mont_array(pub), log_n, w) == why
    # a second rank layout of the same proof: sharded over the ranks the blowup allows, on this one device
    comm = pkg.Comm.local(ctx, 2)
    assert np.array_equal(pkg.prove_sharded(comm, pkg.FriConfig(**fri_kw), _gpu_cfgs(pkg, cfgs), dev, pkg.from_mont_array(pub)).words, gproof.words)
    comm.close()
    if port_seconds <= 60 or os.environ.get("LSP_FULL_PARITY"):
        assert np.array_equal(cport.prove_limbs(ofri, tr, n, w, cfgs, pub), gproof.words), name
    ctx.close()


def test_cfg5_standin_cbor_file_through_lsp_prove_is_bit_exact(pkg, tmp_path):
    """BASELINE configs[4] (the stripped zkevm.bin): its stand-in is a 6+6-column `RawPermutationTrace` CBOR file of 2^19
    rows (the `mxp` shape of bench.log:2-13, ~390 MB).  The C++ `main` (lsp_prove: file -> device witness -> prove ->
    device verify) must write the proof the Python mirror produces from the same bytes, and that proof must be the CPU
    port's, word for word, on the 14-column trace."""
    import os
    import subprocess
    import numpy as np
    from oracle import cport
    from oracle.poseidon2 import Poseidon2Params
    from tests.proofs import fast_cbor_permutation_trace
    from tests.test_gpu_main_driver import EXE, _draws
    if os.environ.get("LSP_SKIP_BIG"):
        pytest.skip("LSP_SKIP_BIG set")
    log_n, c, seed = int(os.environ.get("LSP_CFG5_LOG_N", "19")), 6, 20260
    n = 1 << log_n
    rng = np.random.default_rng(5)
    a = rng.integers(0, 256, size=(n, c, 32), dtype=np.uint8)
    a[:, :, 0] &= 0x0F                                           # below 2^252 < r: canonical values
    be = np.ascontiguousarray(np.concatenate([a, a[rng.permutation(n)]], axis=1)).reshape(-1)
    path = tmp_path / "mxp.bin"
    path.write_bytes(fast_cbor_permutation_trace(be, n, c, "mxp"))
    out = tmp_path / "proof.bin"
    r = subprocess.run([str(EXE), "--permutation", str(path), "--seed", str(seed), "--out", str(out)], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "proof accepted" in r.stdout and f"{n} rows x 14 columns" in r.stdout
    words = np.frombuffer(out.read_bytes(), dtype=np.uint64)
    path.unlink()
    alpha, delta, consts = _draws(seed)
    p = Poseidon2Params(sbox_d=5, rounds_f=8, rounds_p=22, ext_initial=[consts[3 * i:3 * i + 3] for i in range(4)],
                        ext_terminal=[consts[12 + 3 * i:15 + 3 * i] for i in range(4)], internal=consts[24:], internal_diag_m1=(1, 1, 2))
    ctx = pkg.Context(0)
    ctx.set_poseidon2(5, 8, 22, p.flat_constants(), p.internal_diag_m1)
    pub = pkg.to_mont_array([alpha, delta])
    trace = ctx.permutation_trace_be(be, n, c, pub)
    cfgs = [OA.AirPermutationConfig.standard(c)]
    mine = pkg.prove(ctx, pkg.FriConfig(), _gpu_cfgs(pkg, cfgs), trace, [alpha, delta])
    assert np.array_equal(words, mine.words)
    cport.set_poseidon2(p)
    cport.set_threads(0)
    tr = trace.download_array()
    assert np.array_equal(cport.prove_limbs(OS.FriConfig(), tr, n, 14, cfgs, pub), mine.words)
    ctx.close()


@pytest.mark.parametrize("d,rf,rp,diag", [(3, 8, 22, (1, 1, 2)), (7, 8, 22, (1, 1, 2)), (11, 8, 22, (1, 1, 2)), (17, 8, 22, (1, 1, 2)),
                                          (5, 6, 13, (3, 5, 7)), (17, 4, 0, (1, 1, 2))])
def test_prove_with_other_poseidon2_parameters(pkg, d, rf, rp, diag):
    """Nothing fork-specific is baked in: S-box degree, round counts and the internal
    diagonal are run-time parameters of both permutation shapes (one thread / three lanes per permutation),
    so whole proofs must stay bit-identical to the oracle's for every supported choice."""
    from oracle.poseidon2 import Poseidon2Params
    p = Poseidon2Params.from_seed(1000 + d, sbox_d=d, rounds_f=rf, rounds_p=rp)
    p.internal_diag_m1 = diag
    ctx = pkg.Context(0)
    ctx.set_poseidon2(d, rf, rp, p.flat_constants(), diag)
    cfgs, trace, publics = _perm_instance(5, 2, 60 + d)
    fri = dict(log_blowup=2, log_final_poly_len=0, num_queries=6, proof_of_work_bits=3)
    gd, _ = pkg.prove(ctx, pkg.FriConfig(**fri), _gpu_cfgs(pkg, cfgs), trace, publics).to_dict()
    assert gd == OS.prove(p, OS.FriConfig(**fri), cfgs, trace, publics)
    OS.verify(p, OS.FriConfig(**fri), cfgs, gd, publics)
    # both permutation shapes on the same states: thread-per-permutation probe vs the tree the tri-lane kernel built
    rng = F.SplitMix64(d)
    states = [[rng.next_fr(), rng.next_fr(), 0] for _ in range(64)]
    assert [s[0] for s in ctx.permute(states)] == [OP.compress(p, s[0], s[1]) for s in states]
    ctx.close()


@pytest.mark.parametrize("log_n,w,bits", [(3, 3, 2), (6, 5, 1), (8, 2, 3)])
def test_pcs_open_pieces_eval_at_and_reduce_openings(pkg, gctx, p2params, log_n, w, bits):
    """`Pcs::open` from its pieces (lsp_eval_at, lsp_reduce_openings) for a trait-level drop-in that keeps its own
    challenger: opened values and the FRI input vector equal the oracle's `pcs_open` on a trace-like commit (two points) and
    a quotient-like commit (two one-column matrices on the split, shifted domains)."""
    from oracle.challenger import HashChallenger
    rng = F.SplitMix64(1000 + log_n)
    n = 1 << log_n
    fri = OS.FriConfig(log_blowup=bits, log_final_poly_len=0, num_queries=2, proof_of_work_bits=0)
    mat = [[rng.next_fr() for _ in range(w)] for _ in range(n)]
    chunks = [[[rng.next_fr()] for _ in range(n)] for _ in range(2)]
    dom = OS.Domain(log_n, 1)
    qdoms = dom.create_disjoint_domain(2 * n).split_domains(2)
    _, tree = OS.pcs_commit(p2params, fri, [(dom, mat)])
    _, qtree = OS.pcs_commit(p2params, fri, list(zip(qdoms, chunks)))
    z = rng.next_fr()
    z2 = z * dom.gen() % F.R_MOD
    ch = HashChallenger(p2params, [])
    ch.observe(7)
    dbg = {}
    opened, _ = OS.pcs_open(p2params, fri, [(tree, [[z, z2]]), (qtree, [[z], [z]])], ch, dbg)
    dft = pkg.GpuDft(gctx)
    lde_t, co_t = dft.coset_lde_batch(gctx.upload(mat), bits, F.GENERATOR, want_coeffs=True)
    assert pkg.eval_at(gctx, co_t, z) == opened[0][0][0] and pkg.eval_at(gctx, co_t, z2) == opened[0][0][1]
    entries = [(lde_t, z, opened[0][0][0]), (lde_t, z2, opened[0][0][1])]
    for c, qd in enumerate(qdoms):
        lde_c, co_c = dft.coset_lde_batch(gctx.upload(chunks[c]), bits, F.GENERATOR * F.inv(qd.shift) % F.R_MOD, want_coeffs=True)
        assert pkg.eval_at(gctx, co_c, z * F.inv(qd.shift) % F.R_MOD) == opened[1][c][0]
        entries.append((lde_c, z, opened[1][c][0]))
    got = pkg.reduce_openings(gctx, entries, dbg["fri_alpha"]).rows()
    assert [r[0] for r in got] == dbg["fri_input"]
