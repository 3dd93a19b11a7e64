"""The device verifier (`lsp_verify_air`: `verify(&config, &air, &mut challenger, &proof, &publics)`,
bin/src/main.rs:88-96) and `Mmcs::verify_batch` against the oracle's verifiers: same verdict on honest proofs,
and on tampered ones the SAME rejection reason the CPU verifier reports, region by region of the proof."""
import numpy as np
import pytest

from oracle import air as OA
from oracle import field as F
from oracle import merkle as OM
from oracle import stark as OS
from oracle import trace as OT

pytestmark = pytest.mark.gpu


def _gpu_cfgs(pkg, cfgs):
    out = []
    for c in cfgs:
        if isinstance(c, OA.AirLookupConfig):
            out.append(pkg.AirLookupConfig(c.a_columns_ids, c.b_columns_ids, c.a_filter_id, c.b_filter_id, c.a_inverses_id,
                                           c.b_inverses_id, c.occurrences_id, c.check_id))
        else:
            out.append(pkg.AirPermutationConfig(c.a_columns_ids, c.b_columns_ids, c.b_inverse_id, c.check_id))
    return out


def _perm_case(pkg, gctx, cport, log_n, cols, fri_kw, seed):
    """A proof of `len(cols)` permutation arguments side by side; returns everything both verifiers need."""
    n = 1 << log_n
    rng = F.SplitMix64(seed)
    alpha, delta = rng.next_fr(), rng.next_fr()
    pm = [OT.synthetic_permutation_input(seed + 10 + i, c, n) for i, c in enumerate(cols)]
    cfgs, trace = OT.build_trace(pm, alpha, delta, [])
    g = _gpu_cfgs(pkg, cfgs)
    gproof = pkg.prove(gctx, pkg.FriConfig(**fri_kw), g, trace, [alpha, delta])
    return cfgs, g, gproof, [alpha, delta], pkg.to_mont_array([alpha, delta])


@pytest.mark.parametrize("log_n,cols,fri_kw", [
    (1, [1], dict(log_blowup=1, log_final_poly_len=0, num_queries=3, proof_of_work_bits=0)),
    (4, [3], dict(log_blowup=3, log_final_poly_len=0, num_queries=33, proof_of_work_bits=0)),
    (6, [2, 1], dict(log_blowup=2, log_final_poly_len=2, num_queries=11, proof_of_work_bits=4)),
    (9, [5], dict(log_blowup=1, log_final_poly_len=1, num_queries=40, proof_of_work_bits=1)),
    (11, [1, 2, 3], dict(log_blowup=3, log_final_poly_len=3, num_queries=7, proof_of_work_bits=0)),
])
def test_honest_permutation_proofs_are_accepted(pkg, gctx, p2params, log_n, cols, fri_kw):
    from oracle import cport
    cport.set_poseidon2(p2params)
    cfgs, g, gproof, publics, pub = _perm_case(pkg, gctx, cport, log_n, cols, fri_kw, 600 + log_n)
    ofri = OS.FriConfig(**fri_kw)
    assert cport.verify_limbs(ofri, log_n, gproof.width, cfgs, pub, gproof.words) == 0
    tm = {}
    assert pkg.verify(gctx, pkg.FriConfig(**fri_kw), g, gproof, publics, timing=tm) is None
    assert tm["device_ms"] > 0
    # a flat word array works as well (what a host holding only the bytes has)
    assert pkg.verify_code(gctx, pkg.FriConfig(**fri_kw), g, gproof.words.copy(), publics, log_n, gproof.width) == 0


def _regions(log_n, w, q, fri):
    """Named element ranges of the flat proof (host/prover.cu layout)."""
    log_l = log_n + fri["log_blowup"]
    rounds = log_n - fri["log_final_poly_len"]
    f = 1 << (fri["log_blowup"] + fri["log_final_poly_len"])
    reg, pos = {}, 0

    def add(name, k):
        nonlocal pos
        if k:
            reg[name] = (pos, pos + k)
        pos += k

    add("trace_commit", 1), add("quotient_commit", 1), add("trace_local", w), add("trace_next", w), add("chunks", q)
    add("fri_commits", rounds), add("final_poly", f), add("pow_witness", 1)
    for qi in range(fri["num_queries"]):
        add(f"q{qi}.index", 1), add(f"q{qi}.trace_row", w), add(f"q{qi}.trace_path", log_l)
        add(f"q{qi}.quot_row", q), add(f"q{qi}.quot_path", log_l)
        for r in range(rounds):
            add(f"q{qi}.r{r}.sibling", 1), add(f"q{qi}.r{r}.path", log_l - 1 - r)
    return reg, pos


@pytest.mark.parametrize("pow_bits", [0, 3])
def test_rejection_reasons_match_the_c_port_region_by_region(pkg, gctx, p2params, pow_bits):
    """Flip one bit in every region of the proof (first and last element of each, three queries' worth) and in 80
    random words: the device verifier must return exactly the code the C port's verifier returns, never 0."""
    from oracle import cport
    cport.set_poseidon2(p2params)
    log_n, fri_kw = 6, dict(log_blowup=2, log_final_poly_len=1, num_queries=9, proof_of_work_bits=pow_bits)
    cfgs, g, gproof, publics, pub = _perm_case(pkg, gctx, cport, log_n, [2], fri_kw, 7000 + pow_bits)
    ofri, gfri, w = OS.FriConfig(**fri_kw), pkg.FriConfig(**fri_kw), gproof.width
    reg, n_elems = _regions(log_n, w, 2, fri_kw)
    assert n_elems * 4 == gproof.words.size
    assert pkg.verify_code(gctx, gfri, g, gproof, publics) == 0
    seen = set()

    def check(word, bit, label):
        bad = gproof.words.copy()
        bad[word] ^= np.uint64(1 << bit)
        want = cport.verify_limbs(ofri, log_n, w, cfgs, pub, bad)
        got = pkg.verify_code(gctx, gfri, g, bad, publics, log_n, w)
        assert got == want and got != 0, (label, word, bit, got, want)
        seen.add(got)

    for name, (lo, hi) in reg.items():
        if name.startswith("q") and name[1].isdigit() and int(name[1:].split(".")[0]) not in (0, 4, 8):
            continue
        for e in {lo, hi - 1}:
            check(4 * e, 0, name)
    rng = np.random.default_rng(99 + pow_bits)
    for _ in range(80):
        word = int(rng.integers(0, gproof.words.size))
        word -= word % 4 == 3            # keep the tampered element below r: leave the top limb alone
        check(word, int(rng.integers(0, 64)), "random")
    # Every reason a single flipped bit can produce shows up.  (A flip in an opened value or the final polynomial is
    # caught by a query before the final-polynomial / out-of-domain checks are reached: those two reasons are
    # exercised by test_zero_round_proof_is_rejected_like_the_port and test_wrong_statement_is_rejected.)
    assert seen >= ({1, 2, 3, 4} | ({6} if pow_bits else set())), seen


def test_wrong_statement_is_rejected(pkg, gctx, p2params):
    """The right proof for the wrong publics / FRI parameters / AIR."""
    from oracle import cport
    cport.set_poseidon2(p2params)
    log_n, fri_kw = 5, dict(log_blowup=2, log_final_poly_len=0, num_queries=6, proof_of_work_bits=2)
    cfgs, g, gproof, publics, pub = _perm_case(pkg, gctx, cport, log_n, [2], fri_kw, 31337)
    gfri, w = pkg.FriConfig(**fri_kw), gproof.width
    with pytest.raises(pkg.VerificationError) as e:
        pkg.verify(gctx, gfri, g, gproof, [publics[0], (publics[1] + 1) % F.R_MOD])
    pub2 = pkg.to_mont_array([publics[0], (publics[1] + 1) % F.R_MOD])
    assert e.value.code == cport.verify_limbs(OS.FriConfig(**fri_kw), log_n, w, cfgs, pub2, gproof.words)
    # one query fewer / one more blowup bit: another proof shape altogether
    for other in (dict(fri_kw, num_queries=5), dict(fri_kw, log_blowup=3), dict(fri_kw, log_final_poly_len=1)):
        assert pkg.verify_code(gctx, pkg.FriConfig(**other), g, gproof, publics) == 1
        assert cport.verify_limbs(OS.FriConfig(**other), log_n, w, cfgs, pub, gproof.words) == 1
    # more proof-of-work bits than the prover ground for: same shape, witness (almost surely) fails
    hard = dict(fri_kw, proof_of_work_bits=30)
    assert pkg.verify_code(gctx, pkg.FriConfig(**hard), g, gproof, publics) == \
        cport.verify_limbs(OS.FriConfig(**hard), log_n, w, cfgs, pub, gproof.words) == 6
    # the columns of a and b swapped in the AIR: constraints no longer hold at zeta
    c0 = cfgs[0]
    swapped = [OA.AirPermutationConfig(list(reversed(c0.a_columns_ids)), c0.b_columns_ids, c0.b_inverse_id, c0.check_id)]
    assert pkg.verify_code(gctx, gfri, _gpu_cfgs(pkg, swapped), gproof, publics) == \
        cport.verify_limbs(OS.FriConfig(**fri_kw), log_n, w, swapped, pub, gproof.words) == 7
    # a truncated buffer
    assert pkg.verify_code(gctx, gfri, g, gproof.words[:-4].copy(), publics, log_n, w) == 1


@pytest.mark.parametrize("log_n,lookups,perms,fri_kw", [
    (3, [(1, 1, 0)], [], dict(log_blowup=2, log_final_poly_len=0, num_queries=5, proof_of_work_bits=0)),
    (6, [(3, 1, 0), (2, 2, 9)], [2], dict(log_blowup=3, log_final_poly_len=1, num_queries=7, proof_of_work_bits=3)),
])
def test_lookup_air_proofs(pkg, gctx, p2params, log_n, lookups, perms, fri_kw):
    """q = 4 quotient chunks: the recombination sum_i zp_i(zeta) chunk_i(zeta) has 12 cross factors."""
    from oracle import cport
    cport.set_poseidon2(p2params)
    n = 1 << log_n
    rng = F.SplitMix64(555 + log_n)
    alpha, delta = rng.next_fr(), rng.next_fr()
    lk = [OT.synthetic_lookup_input(17 + 3 * i, nc, nt, n, disabled_every=de) for i, (nc, nt, de) in enumerate(lookups)]
    pm = [OT.synthetic_permutation_input(170 + i, c, n) for i, c in enumerate(perms)]
    cfgs, trace = OT.build_trace(pm, alpha, delta, lk)
    g, gfri, ofri = _gpu_cfgs(pkg, cfgs), pkg.FriConfig(**fri_kw), OS.FriConfig(**fri_kw)
    gproof = pkg.prove(gctx, gfri, g, trace, [alpha, delta])
    pub, w = pkg.to_mont_array([alpha, delta]), gproof.width
    assert gproof.log_q == 2
    pkg.verify(gctx, gfri, g, gproof, [alpha, delta])
    gd, _ = gproof.to_dict()
    OS.verify(p2params, ofri, cfgs, gd, [alpha, delta])
    for e in (2, 2 + w - 1, 2 + w, 2 + 2 * w, 2 + 2 * w + 3):       # opened local/next values and the first/last chunk
        bad = gproof.words.copy()
        bad[4 * e + 1] ^= np.uint64(1 << 9)
        want = cport.verify_limbs(ofri, log_n, w, cfgs, pub, bad)
        assert want != 0 and pkg.verify_code(gctx, gfri, g, bad, [alpha, delta], log_n, w) == want


def test_zero_round_fri_is_refused_by_prover_and_verifier(pkg, gctx, p2params):
    """log_final_poly_len == log2(height): no commit-phase round, so the reduced opening never enters the fold chain and
    the low-degree test would be vacuous (the pinned verifier rejects every honest such proof with FinalPolyMismatch --
    oracle tests keep that pinned -- and a forged all-zero final polynomial would pass it).  The library refuses the
    configuration on both sides with LSP_ERR_PARAM; so does a proof_of_work_bits above the 32 that sample_bits can test."""
    from oracle import cport
    cport.set_poseidon2(p2params)
    ok_kw = dict(log_blowup=2, log_final_poly_len=2, num_queries=4, proof_of_work_bits=0)
    rng = F.SplitMix64(4242)
    alpha, delta = rng.next_fr(), rng.next_fr()
    cfgs, trace = OT.build_trace([OT.synthetic_permutation_input(4252, 1, 8)], alpha, delta, [])
    g, publics, pub = _gpu_cfgs(pkg, cfgs), [alpha, delta], pkg.to_mont_array([alpha, delta])
    gproof = pkg.prove(gctx, pkg.FriConfig(**ok_kw), g, trace, publics)
    pkg.verify(gctx, pkg.FriConfig(**ok_kw), g, gproof, publics)
    for bad_kw, msg in ((dict(ok_kw, log_final_poly_len=3), "no commit-phase round"), (dict(ok_kw, log_final_poly_len=4), "no commit-phase round"),
                        (dict(ok_kw, proof_of_work_bits=33), "proof_of_work_bits")):
        with pytest.raises(pkg.BackendError, match=msg):
            pkg.prove(gctx, pkg.FriConfig(**bad_kw), g, trace, publics)
        with pytest.raises(pkg.BackendError, match=msg):
            pkg.verify_code(gctx, pkg.FriConfig(**bad_kw), g, gproof.words, publics, log_n=3, width=gproof.width)
    # the CPU port's (release-build) behaviour that the refusal replaces stays pinned: its own zero-round proof fails at the final polynomial
    zfri = OS.FriConfig(**dict(ok_kw, log_final_poly_len=3))
    zwords = cport.prove(zfri, cfgs, trace, publics)
    assert cport.verify_limbs(zfri, 3, gproof.width, cfgs, pub, zwords) == 5


@pytest.mark.parametrize("d", [3, 7, 11, 17])
def test_other_sbox_degrees(pkg, d):
    """The unpinned Poseidon2 S-box degree is a parameter of the verifier as well."""
    from oracle import cport
    from oracle.poseidon2 import Poseidon2Params
    p = Poseidon2Params.from_seed(0xD00 + d, sbox_d=d)
    cport.set_poseidon2(p)
    ctx = pkg.Context(0)
    ctx.set_poseidon2(p.sbox_d, p.rounds_f, p.rounds_p, p.flat_constants(), p.internal_diag_m1)
    fri_kw = dict(log_blowup=2, log_final_poly_len=0, num_queries=5, proof_of_work_bits=2)
    cfgs, g, gproof, publics, pub = _perm_case(pkg, ctx, cport, 4, [2], fri_kw, 800 + d)
    assert cport.verify_limbs(OS.FriConfig(**fri_kw), 4, gproof.width, cfgs, pub, gproof.words) == 0
    pkg.verify(ctx, pkg.FriConfig(**fri_kw), g, gproof, publics)
    bad = gproof.words.copy()
    bad[-1] ^= np.uint64(2)
    assert pkg.verify_code(ctx, pkg.FriConfig(**fri_kw), g, bad, publics, 4, gproof.width) == \
        cport.verify_limbs(OS.FriConfig(**fri_kw), 4, gproof.width, cfgs, pub, bad) != 0
    ctx.close()


@pytest.mark.parametrize("log_h,widths", [(0, [1]), (1, [2]), (5, [3]), (9, [1, 4]), (12, [8])])
def test_mmcs_verify_batch(pkg, gctx, p2params, log_h, widths):
    """`verify_batch` on the device == `oracle/merkle.py:verify_batch` for every kind of damage."""
    h = 1 << log_h
    rng = F.SplitMix64(70 + log_h)
    mats = [[[rng.next_fr() for _ in range(w)] for _ in range(h)] for w in widths]
    mm = pkg.GpuMmcs(gctx)
    root, tree = mm.commit([gctx.upload(m) for m in mats])
    for idx in sorted({0, h // 3, h - 1}):
        rows, proof = mm.open_batch(idx, tree)
        assert OM.verify_batch(p2params, root, h, idx, rows, proof)
        assert mm.verify_batch(root, log_h, idx, rows, proof)
        bad_rows = [list(r) for r in rows]
        bad_rows[-1][-1] = (bad_rows[-1][-1] + 1) % F.R_MOD
        assert not mm.verify_batch(root, log_h, idx, bad_rows, proof)
        assert not mm.verify_batch((root + 1) % F.R_MOD, log_h, idx, rows, proof)
        if log_h:
            for k in {0, log_h - 1}:
                bad_proof = list(proof)
                bad_proof[k] = (bad_proof[k] + 5) % F.R_MOD
                assert not mm.verify_batch(root, log_h, idx, rows, bad_proof)
            assert not mm.verify_batch(root, log_h, idx ^ 1, rows, proof)
    tree.free()


def test_non_canonical_elements_do_not_deserialise(pkg, gctx, p2params):
    """x and x + r are the same field element but not the same encoding: `Bls12_377Fr`'s deserialiser refuses the second,
    so both verifiers report a malformed proof wherever it appears -- even in a slot where the arithmetic would not care."""
    from oracle import cport
    cport.set_poseidon2(p2params)
    log_n, fri_kw = 5, dict(log_blowup=2, log_final_poly_len=0, num_queries=5, proof_of_work_bits=0)
    cfgs, g, gproof, publics, pub = _perm_case(pkg, gctx, cport, log_n, [2], fri_kw, 2718)
    reg, _ = _regions(log_n, gproof.width, 2, fri_kw)
    vals = pkg.from_mont_array(gproof.words.reshape(-1, 4))       # canonical integers of the Montgomery residues' VALUES
    for name in ("trace_local", "chunks", "final_poly", "q0.trace_row", "q4.quot_path", "q2.r1.sibling"):
        e = reg[name][0]
        limbs = gproof.words.reshape(-1, 4)[e]
        stored = sum(int(limbs[k]) << (64 * k) for k in range(4))          # the Montgomery representative as stored
        if stored + F.R_MOD >= 1 << 256:
            continue
        bad = gproof.words.copy()
        for k in range(4):
            bad[4 * e + k] = np.uint64(((stored + F.R_MOD) >> (64 * k)) & (2 ** 64 - 1))
        assert cport.verify_limbs(OS.FriConfig(**fri_kw), log_n, gproof.width, cfgs, pub, bad) == 1, name
        assert pkg.verify_code(gctx, pkg.FriConfig(**fri_kw), g, bad, publics, log_n, gproof.width) == 1, name
    del vals
