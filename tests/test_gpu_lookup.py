"""LogUp lookup AIR: `LineaAIR::eval_lookup` (air/src/lib.rs:57-114) in the fused
quotient kernel and the full prove with q = 4 quotient chunks, against the oracle.  This is the AIR the
reference's `main` proves at HEAD (bin/src/main.rs:37-43).  Bit-exact."""
import copy

import pytest

from oracle import air as OA
from oracle import dft as OD
from oracle import field as F
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


def _instance(log_n, lookups, perms, seed):
    """lookups: list of (n_cols, n_tables, disabled_every); perms: list of column counts."""
    n = 1 << log_n
    rng = F.SplitMix64(seed)
    alpha, delta = rng.next_fr(), rng.next_fr()
    lk = [OT.synthetic_lookup_input(seed + 3 * i, nc, nt, n, disabled_every=de) for i, (nc, nt, de) in enumerate(lookups)]
    pm = [OT.synthetic_permutation_input(seed + 100 + i, c, n) for i, c in enumerate(perms)]
    cfgs, trace = OT.build_trace(pm, alpha, delta, lk)
    assert OA.check_constraints(cfgs, trace, [alpha, delta])
    return cfgs, trace, [alpha, delta]


def test_lookup_witness_layout_matches_reference_ids():
    """Column ids of `get_air_lookup_config` (trace/src/lookup.rs:178-214) and `AirLookupConfig::width`."""
    c = OA.AirLookupConfig.standard(2, 3, 2)
    assert c.a_columns_ids == [0, 1] and c.b_columns_ids == [[2, 3], [4, 5], [6, 7]]
    assert (c.a_filter_id, c.b_filter_id, c.a_inverses_id) == (8, [9, 10, 11], 12)
    assert (c.b_inverses_id, c.occurrences_id, c.check_id) == ([13, 14, 15], [16, 17, 18], 19)
    assert c.width() == 20 == 2 + 3 * (2 + 3) + 3


@pytest.mark.parametrize("log_n,lookups,perms", [
    (2, [(1, 1, 0)], []),
    (4, [(2, 2, 5)], []),
    (5, [(3, 1, 0), (1, 3, 7)], [2]),
    (7, [(2, 1, 3)], [3, 1]),
])
def test_quotient_values_with_lookups(pkg, gctx, log_n, lookups, perms):
    cfgs, trace, publics = _instance(log_n, lookups, perms, 900 + log_n)
    log_q = OA.log_quotient_degree(cfgs)
    assert log_q == 2
    lde = OD.coset_lde_batch(trace, 3, F.GENERATOR)
    td, qd = OS.Domain(log_n, 1), OS.Domain(log_n + log_q, F.GENERATOR)
    alpha = F.SplitMix64(77).next_fr()
    qv = OS.quotient_values(cfgs, publics, td, qd, OD.bit_reverse_rows(lde[:qd.size()]), alpha)
    got = pkg.quotient_air(gctx, gctx.upload(lde), log_n, _gpu_cfgs(pkg, cfgs), publics, alpha).rows()
    assert got == [[qv[k * 4 + ch] for ch in range(4)] for k in range(1 << log_n)]


@pytest.mark.parametrize("log_n,lookups,perms,fri", [
    (3, [(1, 1, 0)], [], dict(log_blowup=2, log_final_poly_len=0, num_queries=5, proof_of_work_bits=0)),
    (5, [(2, 2, 6)], [2], dict(log_blowup=3, log_final_poly_len=0, num_queries=33, proof_of_work_bits=0)),
    (6, [(3, 1, 0), (2, 2, 9)], [], dict(log_blowup=3, log_final_poly_len=1, num_queries=7, proof_of_work_bits=3)),
])
def test_prove_lookup_air_bit_exact_and_verifies(pkg, gctx, p2params, log_n, lookups, perms, fri):
    cfgs, trace, publics = _instance(log_n, lookups, perms, 40 + log_n)
    ofri = OS.FriConfig(**fri)
    dbg = {}
    oproof = OS.prove(p2params, ofri, cfgs, trace, publics, dbg)
    OS.verify(p2params, ofri, cfgs, oproof, publics)
    gproof = pkg.prove(gctx, pkg.FriConfig(**fri), _gpu_cfgs(pkg, cfgs), trace, publics)
    assert gproof.log_q == 2
    gd, indices = gproof.to_dict()
    assert len(gd["opened_values"]["quotient_chunks"]) == 4
    assert gd == oproof and indices == dbg["query_indices"]
    OS.verify(p2params, ofri, cfgs, gd, publics)
    bad = copy.deepcopy(gd)
    qc = bad["opened_values"]["quotient_chunks"]
    if isinstance(qc[3], list):
        qc[3][0] = (qc[3][0] + 1) % F.R_MOD
    else:
        qc[3] = (qc[3] + 1) % F.R_MOD
    with pytest.raises(OS.VerificationError):
        OS.verify(p2params, ofri, cfgs, bad, publics)


def test_lookup_air_needs_blowup_4(pkg, gctx):
    cfgs, trace, publics = _instance(3, [(1, 1, 0)], [], 5)
    with pytest.raises(pkg.BackendError, match="quotient degree"):
        pkg.prove(gctx, pkg.FriConfig(log_blowup=1), _gpu_cfgs(pkg, cfgs), trace, publics)


def test_lookup_configs_must_come_first(pkg, gctx):
    cfgs, trace, publics = _instance(3, [(1, 1, 0)], [1], 6)
    with pytest.raises(pkg.BackendError, match="lookups before permutations"):
        pkg.prove(gctx, pkg.FriConfig(), list(reversed(_gpu_cfgs(pkg, cfgs))), trace, publics)


def _lookup_rows(a, b, af, bf):
    n = len(a[0])
    cols = list(a) + [c for t in b for c in t] + [af] + list(bf)
    return [x for i in range(n) for x in (col[i] for col in cols)]


@pytest.mark.parametrize("log_n,n_cols,n_tables,disabled", [(0, 1, 1, 0), (3, 1, 1, 0), (5, 2, 3, 4), (9, 3, 2, 7), (12, 1, 2, 0)])
def test_lookup_witness_on_device(pkg, gctx, log_n, n_cols, n_tables, disabled):
    """lsp_lookup_trace == RawLookupTrace::get_trace (trace/src/lookup.rs:46-176), column for column."""
    n = 1 << log_n
    rng = F.SplitMix64(300 + log_n)
    alpha, delta = rng.next_fr(), rng.next_fr()
    a, b, af, bf = OT.synthetic_lookup_input(50 + log_n, n_cols, n_tables, n, disabled_every=disabled)
    if n >= 8:
        # the same row in two tables, and a DISABLED earlier holder: the count must go to the first ENABLED holder
        for k in range(n_cols):
            b[-1][k][1] = b[0][k][2]
        bf[0][0] = 0
    cfg, cols = OT.lookup_columns(a, b, af, bf, alpha, delta)
    pub = pkg.to_mont_array([alpha, delta])
    got = gctx.lookup_trace(pkg.to_mont_array(_lookup_rows(a, b, af, bf)), n, n_cols, n_tables, n_cols, pub)
    assert got.width == cfg.width()
    assert got.rows() == OT.row_major(cols)


def test_lookup_witness_rejects_missing_value(pkg, gctx):
    rng = F.SplitMix64(9)
    alpha, delta = rng.next_fr(), rng.next_fr()
    a, b, af, bf = OT.synthetic_lookup_input(60, 2, 1, 16)
    a[0][3] = 424242
    pub = pkg.to_mont_array([alpha, delta])
    with pytest.raises(pkg.BackendError, match="check column should be 0"):
        gctx.lookup_trace(pkg.to_mont_array(_lookup_rows(a, b, af, bf)), 16, 2, 1, 2, pub)
    af[3] = 0   # the filter switches the offending row off
    gctx.lookup_trace(pkg.to_mont_array(_lookup_rows(a, b, af, bf)), 16, 2, 1, 2, pub)


def test_main_flow_cbor_files_to_proof(pkg, gctx, p2params):
    """What the reference's `main` does at HEAD (bin/src/main.rs:35-86): read a lookup file (and here a
    permutation file too), `push_traces` (lookups first), `LineaAIR::new(cfgs)`, `prove` -- all on the device
    from the CBOR bytes; trace and proof equal the oracle's."""
    n = 32
    rng = F.SplitMix64(12)
    alpha, delta = rng.next_fr(), rng.next_fr()
    lk = OT.synthetic_lookup_input(70, 2, 2, n)
    # filters partly omitted in the file -> `read_file` defaults them to one: a_filter = [0, 1, 1, 1, ...]
    lk_blob = OT.encode_raw_lookup_trace(lk[0], lk[1], [0, 1, 1], [lk[3][0]], "lookup_0")
    pa, pb = OT.synthetic_permutation_input(71, 3, n)
    pm_blob = OT.encode_raw_permutation_trace(pa, pb, "perm_0")
    pub = pkg.to_mont_array([alpha, delta])
    be, rows, na, nt, nb, _ = pkg.read_raw_lookup_trace(lk_blob)
    t_lk = gctx.lookup_trace(be, rows, na, nt, nb, pub)
    be, rows, nc, _ = pkg.read_raw_permutation_trace(pm_blob)
    t_pm = gctx.permutation_trace_be(be, rows, nc, pub)
    trace_dev = gctx.hconcat([t_lk, t_pm])
    dl = OT.decode_raw_lookup_trace(lk_blob)
    cfgs, trace = OT.build_trace([(pa, pb)], alpha, delta, [dl[:4]])
    assert trace_dev.rows() == trace
    fri = dict(log_blowup=3, log_final_poly_len=0, num_queries=6, proof_of_work_bits=0)
    gd, _ = pkg.prove(gctx, pkg.FriConfig(**fri), _gpu_cfgs(pkg, cfgs), trace_dev, [alpha, delta]).to_dict()
    oproof = OS.prove(p2params, OS.FriConfig(**fri), cfgs, trace, [alpha, delta])
    assert gd == oproof
    OS.verify(p2params, OS.FriConfig(**fri), cfgs, gd, [alpha, delta])


@pytest.mark.parametrize("log_n,lookups,perms,world,blowup", [(4, [(1, 1, 0)], [], 2, 2), (5, [(2, 2, 5)], [2], 4, 3), (6, [(1, 2, 0)], [1], 8, 3),
                                                              (10, [(2, 1, 7)], [3], 2, 3),
                                                              (6, [(1, 1, 0)], [1], 8, 2), (7, [(2, 2, 3)], [], 16, 2)])   # more ranks than cosets
def test_sharded_prove_of_lookup_air_equals_single_gpu(pkg, gctx, p2params, log_n, lookups, perms, world, blowup):
    """The sharded prove with 4 quotient chunks (chunk owners spread over the ranks) on the local communicator."""
    import numpy as np
    cfgs, trace, publics = _instance(log_n, lookups, perms, 500 + log_n + world)
    fri = pkg.FriConfig(log_blowup=blowup, num_queries=11)
    g = _gpu_cfgs(pkg, cfgs)
    single = pkg.prove(gctx, fri, g, trace, publics)
    comm = pkg.Comm.local(gctx, world)
    sharded = pkg.prove_sharded(comm, fri, g, trace, publics)
    comm.close()
    assert np.array_equal(single.words, sharded.words)
    if log_n <= 6:
        ofri = OS.FriConfig(log_blowup=blowup, num_queries=11)
        gd, _ = sharded.to_dict()
        assert gd == OS.prove(p2params, ofri, cfgs, trace, publics)


def test_lookup_air_at_2p13_rows_matches_c_port(pkg, gctx, p2params):
    """Beyond the Python oracle's reach: 2^13 rows, device witness == oracle witness, GPU proof == C port's proof,
    accepted by the C verifier; a flipped multiplicity is rejected."""
    import numpy as np
    from oracle import cport
    cport.set_poseidon2(p2params)
    log_n, n = 13, 1 << 13
    rng = F.SplitMix64(2024)
    alpha, delta = rng.next_fr(), rng.next_fr()
    a, b, af, bf = OT.synthetic_lookup_input(88, 2, 2, n, disabled_every=11)
    cfg, cols = OT.lookup_columns(a, b, af, bf, alpha, delta)
    pa, pb = OT.synthetic_permutation_input(89, 2, n)
    cfgs, trace = OT.build_trace([(pa, pb)], alpha, delta, [(a, b, af, bf)])
    pub = pkg.to_mont_array([alpha, delta])
    t_lk = gctx.lookup_trace(pkg.to_mont_array(_lookup_rows(a, b, af, bf)), n, 2, 2, 2, pub)
    ab = pkg.to_mont_array([x for i in range(n) for x in [c[i] for c in pa] + [c[i] for c in pb]])
    dev = gctx.hconcat([t_lk, gctx.permutation_trace(ab, n, 2, pub)])
    limbs = dev.download_array()
    assert np.array_equal(limbs, pkg.to_mont_array([x for r in trace for x in r]))
    fri = dict(log_blowup=3, log_final_poly_len=0, num_queries=33, proof_of_work_bits=0)
    gproof = pkg.prove(gctx, pkg.FriConfig(**fri), _gpu_cfgs(pkg, cfgs), dev, [alpha, delta])
    ofri = OS.FriConfig(**fri)
    w = OA.air_width(cfgs)
    assert cport.verify_limbs(ofri, log_n, w, cfgs, pub, gproof.words) == 0
    assert np.array_equal(cport.prove_limbs(ofri, limbs, n, w, cfgs, pub), gproof.words)
    bad = gproof.words.copy()
    bad[4 * (2 + cfgs[0].occurrences_id[0])] ^= 1
    why = cport.verify_limbs(ofri, log_n, w, cfgs, pub, bad)
    assert why != 0
    # the device verifier: same verdicts, same reason
    pkg.verify(gctx, pkg.FriConfig(**fri), _gpu_cfgs(pkg, cfgs), gproof, [alpha, delta])
    assert pkg.verify_code(gctx, pkg.FriConfig(**fri), _gpu_cfgs(pkg, cfgs), bad, [alpha, delta], log_n, w) == why
