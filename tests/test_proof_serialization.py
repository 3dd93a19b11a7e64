"""Proof serialisation: `lsp_proof_serialize` / `_deserialize` are host-only, so the format is
checked here without a GPU -- against an independent writer that walks the reference-shaped `Proof` dict in field order
(bincode conventions: u64 little-endian lengths, canonical little-endian field elements) -- on proofs made by the CPU port."""
import struct

import numpy as np
import pytest

from oracle import air as OA
from oracle import cport
from oracle import field as F
from oracle import stark as OS
from oracle import trace as OT

MARK = np.uint64(0xFFFFFFFFFFFFFFFF)


def write_reference_shape(d, fri, width, log_q) -> bytes:
    fe = lambda x: int(x).to_bytes(32, "little")
    vec = lambda xs: struct.pack("<Q", len(xs)) + b"".join(fe(x) for x in xs)
    out = b"LSPP" + struct.pack("<7I", 1, fri.log_blowup, fri.log_final_poly_len, fri.num_queries, fri.proof_of_work_bits, width, log_q)
    out += fe(d["commitments"]["trace"]) + fe(d["commitments"]["quotient_chunks"])
    ov = d["opened_values"]
    out += vec(ov["trace_local"]) + vec(ov["trace_next"]) + struct.pack("<Q", len(ov["quotient_chunks"])) + b"".join(vec(c) for c in ov["quotient_chunks"])
    op = d["opening_proof"]
    out += vec(op["commit_phase_commits"]) + struct.pack("<Q", len(op["query_proofs"]))
    for qp in op["query_proofs"]:
        out += struct.pack("<Q", len(qp["input_proof"]))
        for bo in qp["input_proof"]:
            out += struct.pack("<Q", len(bo["opened_values"])) + b"".join(vec(r) for r in bo["opened_values"]) + vec(bo["opening_proof"])
        out += struct.pack("<Q", len(qp["commit_phase_openings"]))
        for st in qp["commit_phase_openings"]:
            out += fe(st["sibling_value"]) + vec(st["opening_proof"])
    out += vec(op["final_poly"]) + fe(op["pow_witness"]) + struct.pack("<Q", d["degree_bits"])
    return out


def port_proof(pkg, p2params, log_n, c, fri_kw, lookups=False):
    cport.set_poseidon2(p2params)
    rng = F.SplitMix64(900 + log_n)
    alpha, delta = rng.next_fr(), rng.next_fr()
    lk = [OT.synthetic_lookup_input(3, 2, 1, 1 << log_n)] if lookups else []
    cfgs, trace = OT.build_trace([OT.synthetic_permutation_input(5, c, 1 << log_n)], alpha, delta, lk)
    words = cport.prove(OS.FriConfig(**fri_kw), cfgs, trace, [alpha, delta])
    return pkg.Proof(words, log_n, OA.air_width(cfgs), OA.log_quotient_degree(cfgs), pkg.FriConfig(**fri_kw))


@pytest.mark.parametrize("log_n,c,lookups,fri_kw", [
    (3, 1, False, dict(log_blowup=1, log_final_poly_len=0, num_queries=2, proof_of_work_bits=0)),
    (5, 3, False, dict(log_blowup=3, log_final_poly_len=2, num_queries=9, proof_of_work_bits=3)),
    (4, 2, True, dict(log_blowup=2, log_final_poly_len=0, num_queries=5, proof_of_work_bits=0)),
])
def test_bytes_follow_the_proof_struct_and_round_trip(pkg, p2params, log_n, c, lookups, fri_kw):
    proof = port_proof(pkg, p2params, log_n, c, fri_kw, lookups)
    blob = proof.serialize()
    d, indices = proof.to_dict()
    assert blob == write_reference_shape(d, proof.fri, proof.width, proof.log_q)
    back = pkg.Proof.deserialize(blob)
    assert (back.log_n, back.width, back.log_q) == (proof.log_n, proof.width, proof.log_q)
    assert vars(back.fri) == vars(proof.fri)
    d2, _ = back.to_dict()
    assert d2 == d
    # the flat arrays agree everywhere but in the per-query index slots, which carry the "not carried" marker
    a, b = proof.words.reshape(-1, 4), back.words.reshape(-1, 4)
    differ = np.where((a != b).any(axis=1))[0]
    assert len(differ) == proof.fri.num_queries and (b[differ] == MARK).all()
    assert [int(a[i][0]) for i in differ] == indices
    assert back.serialize() == blob


def test_malformed_streams_are_rejected(pkg, p2params):
    proof = port_proof(pkg, p2params, 3, 1, dict(log_blowup=1, log_final_poly_len=0, num_queries=2, proof_of_work_bits=0))
    blob = proof.serialize()
    r_le = F.R_MOD.to_bytes(32, "little")
    for bad in (blob[:-1], blob + b"\x00", b"LSPQ" + blob[4:], blob[:4] + struct.pack("<I", 2) + blob[8:],
                blob[:32] + r_le + blob[64:],                                          # a non-canonical element (= r)
                blob[:32 + 64] + struct.pack("<Q", 99) + blob[32 + 64 + 8:],            # a wrong Vec length
                blob[:-8] + struct.pack("<Q", 4)):                                      # degree_bits that does not fit the stream
        with pytest.raises(pkg.BackendError):
            pkg.Proof.deserialize(bad)


def test_absurd_headers_are_refused_at_once(pkg, p2params):
    """A header that promises 2^32 queries (or a 2^32-column trace) must be refused before anything walks the shape it describes
    (found by tools/serialize_asan_harness.cpp: the size pass used to iterate over the promised queries first)."""
    import time
    proof = port_proof(pkg, p2params, 3, 1, dict(log_blowup=1, log_final_poly_len=0, num_queries=2, proof_of_work_bits=0))
    blob = proof.serialize()
    for off, val in ((16, 0xFFFFFFFF), (16, 0), (24, 0xFFFFFFFF), (24, 0), (20, 33), (28, 9)):   # num_queries, width, pow bits, log_q
        bad = blob[:off] + struct.pack("<I", val) + blob[off + 4:]
        t0 = time.perf_counter()
        with pytest.raises(pkg.BackendError):
            pkg.Proof.deserialize(bad)
        assert time.perf_counter() - t0 < 0.5


@pytest.mark.gpu
def test_a_deserialised_proof_verifies_on_the_device(pkg, gctx, p2params):
    fri_kw = dict(log_blowup=2, log_final_poly_len=1, num_queries=6, proof_of_work_bits=2)
    proof = port_proof(pkg, p2params, 5, 2, fri_kw)
    rng = F.SplitMix64(905)
    publics = [rng.next_fr(), rng.next_fr()]
    g = [pkg.AirPermutationConfig(c.a_columns_ids, c.b_columns_ids, c.b_inverse_id, c.check_id) for c in [OA.AirPermutationConfig.standard(2)]]
    pkg.verify(gctx, pkg.FriConfig(**fri_kw), g, proof, publics)
    blob = proof.serialize()
    pkg.verify(gctx, pkg.FriConfig(**fri_kw), g, pkg.Proof.deserialize(blob), publics)        # indices filled in by the verifier
    tampered = bytearray(blob)
    tampered[32 + 7] ^= 1                                                                      # inside the trace commitment
    assert pkg.verify_code(gctx, pkg.FriConfig(**fri_kw), g, pkg.Proof.deserialize(bytes(tampered)), publics) != 0
    # a flat proof that DOES carry indices is still held to the sampled ones
    wrong = proof.words.copy().reshape(-1, 4)
    first_index_slot = np.where((pkg.Proof.deserialize(blob).words.reshape(-1, 4) == MARK).all(axis=1))[0][0]
    wrong[first_index_slot][0] ^= np.uint64(1)
    assert pkg.verify_code(gctx, pkg.FriConfig(**fri_kw), g, wrong.reshape(-1), publics, log_n=5, width=proof.width) == 1
