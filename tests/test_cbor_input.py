"""Input wire format: `RawPermutationTrace` as CBOR, parsed by the library's
host-only entry points (no GPU).  The encoder is the oracle's (cbor2, serde's `[u8;32]`-as-array encoding)."""
import cbor2
import numpy as np
import pytest

from oracle import field as F
from oracle import trace as OT


def _be(cols_a, cols_b, rows):
    c = len(cols_a)
    out = np.zeros((rows, 2 * c, 32), dtype=np.uint8)
    for j, col in enumerate(list(cols_a) + list(cols_b)):
        for i, x in enumerate(col):
            out[i, j] = np.frombuffer(int(x).to_bytes(32, "big"), dtype=np.uint8)
    return out.reshape(-1)


def test_decode_matches_encoder(pkg):
    a, b = OT.synthetic_permutation_input(3, 3, 16)
    a[0][5] = (1 << 256) - 1          # >= r: bytes must come through untouched (reduction happens on the device)
    blob = OT.encode_raw_permutation_trace(a, b, "mxp")
    be, rows, nc, name = pkg.read_raw_permutation_trace(blob)
    assert (rows, nc, name) == (16, 3, "mxp")
    assert np.array_equal(be, _be(a, b, 16))


def test_ragged_columns_are_zero_padded(pkg):
    """`resize` (trace/src/permutation.rs:134-142): short columns grow to the tallest with zeros."""
    a = [[1, 2, 3, 4], [5, 6]]
    b = [[7], [8, 9, 10]]
    be, rows, nc, _ = pkg.read_raw_permutation_trace(OT.encode_raw_permutation_trace(a, b, "r"))
    assert (rows, nc) == (4, 2)
    assert np.array_equal(be, _be([a[0], a[1] + [0, 0]], [b[0] + [0, 0, 0], b[1] + [0]], 4))


def test_byte_strings_indefinite_lengths_and_unknown_keys(pkg):
    e = lambda x: int(x).to_bytes(32, "big")
    plain = cbor2.dumps({"a": [[e(1), e(2)]], "b": [[e(2), e(1)]], "name": "x", "extra": [1, {"k": 2}]})
    be, rows, nc, name = pkg.read_raw_permutation_trace(plain)
    assert (rows, nc, name) == (2, 1, "x") and np.array_equal(be, _be([[1, 2]], [[2, 1]], 2))
    # hand-rolled indefinite-length encoding: map(*) { "a": [* [* elem elem ] ], "b": ..., "name": "y" }
    elem = lambda x: bytes([0x9f]) + b"".join(bytes([v]) if v < 24 else bytes([0x18, v]) for v in e(x)) + b"\xff"
    col = lambda xs: b"\x9f" + b"".join(elem(x) for x in xs) + b"\xff"
    blob = (b"\xbf" + cbor2.dumps("a") + b"\x9f" + col([300, 4]) + b"\xff" + cbor2.dumps("b") + b"\x9f" + col([4, 300]) + b"\xff"
            + cbor2.dumps("name") + cbor2.dumps("y") + b"\xff")
    be, rows, nc, name = pkg.read_raw_permutation_trace(blob)
    assert (rows, nc, name) == (2, 1, "y") and np.array_equal(be, _be([[300, 4]], [[4, 300]], 2))


@pytest.mark.parametrize("mutate", ["truncate", "short_elem", "big_byte", "no_b", "mismatched_cols", "not_cbor"])
def test_malformed_input_is_rejected(pkg, mutate):
    a, b = OT.synthetic_permutation_input(4, 2, 4)
    obj = {"a": [[list(int(x).to_bytes(32, "big")) for x in col] for col in a],
           "b": [[list(int(x).to_bytes(32, "big")) for x in col] for col in b], "name": "t"}
    if mutate == "short_elem":
        obj["a"][0][1] = obj["a"][0][1][:31]
    elif mutate == "big_byte":
        obj["b"][1][2][7] = 256
    elif mutate == "no_b":
        del obj["b"]
    elif mutate == "mismatched_cols":
        obj["b"] = obj["b"][:1]
    blob = cbor2.dumps(obj)
    if mutate == "truncate":
        blob = blob[:len(blob) // 2]
    elif mutate == "not_cbor":
        blob = b"\x00\x01\x02"
    with pytest.raises(pkg.BackendError):
        pkg.read_raw_permutation_trace(blob)


@pytest.mark.gpu
def test_cbor_file_to_proof(pkg, gctx, p2params):
    """BASELINE configs[4] stand-in (the zkevm.bin blob is stripped from the reference): a 6+6-column
    `RawPermutationTrace` CBOR file -> device witness -> prove; trace and proof equal the oracle's."""
    from oracle import air as OA
    from oracle import stark as OS
    n, c = 64, 6
    rng = F.SplitMix64(77)
    alpha, delta = rng.next_fr(), rng.next_fr()
    a, b = OT.synthetic_permutation_input(8, c, n)
    # a few non-canonical encodings of the same row in a and b: `from_be_bytes_mod_order` reduces them
    k = next(i for i in range(n) if all(b[t][i] == a[t][3] for t in range(c)))
    for j in range(c):
        v = a[j][3]
        a[j][3] = v + F.R_MOD
        b[j][k] = v + 2 * F.R_MOD
    blob = OT.encode_raw_permutation_trace(a, b, "mxp")
    be, rows, nc, _ = pkg.read_raw_permutation_trace(blob)
    pub = pkg.to_mont_array([alpha, delta])
    dev = gctx.permutation_trace_be(be, rows, nc, pub)
    da, db, _ = OT.decode_raw_permutation_trace(blob)
    cfgs, trace = OT.build_trace([(da, db)], alpha, delta)
    assert dev.rows() == trace
    fri = dict(log_blowup=3, log_final_poly_len=0, num_queries=9, proof_of_work_bits=0)
    g = [pkg.AirPermutationConfig(x.a_columns_ids, x.b_columns_ids, x.b_inverse_id, x.check_id) for x in cfgs]
    gd, _ = pkg.prove(gctx, pkg.FriConfig(**fri), g, dev, [alpha, delta]).to_dict()
    assert gd == OS.prove(p2params, OS.FriConfig(**fri), cfgs, trace, [alpha, delta])
    assert OA.check_constraints(cfgs, trace, [alpha, delta])


def _be_lookup(a, b, af, bf, rows):
    cols = list(a) + [c for t in b for c in t] + [af] + list(bf)
    out = np.zeros((rows, len(cols), 32), dtype=np.uint8)
    for j, col in enumerate(cols):
        for i, x in enumerate(col):
            out[i, j] = np.frombuffer(int(x).to_bytes(32, "big"), dtype=np.uint8)
    return out.reshape(-1)


def test_lookup_decode_matches_oracle(pkg):
    a, b, af, bf = OT.synthetic_lookup_input(21, 2, 3, 16, disabled_every=4)
    blob = OT.encode_raw_lookup_trace(a, b, af, bf, "lookup_0")
    be, rows, na, nt, nb, name = pkg.read_raw_lookup_trace(blob)
    assert (rows, na, nt, nb, name) == (16, 2, 3, 2, "lookup_0")
    da, db, daf, dbf, _ = OT.decode_raw_lookup_trace(blob)
    assert (da, db, daf, dbf) == (a, b, af, bf)
    assert np.array_equal(be, _be_lookup(a, b, af, bf, 16))


def test_lookup_default_filters_and_padding(pkg):
    """`read_file` fills absent filters with ones up to the guarded column's length (trace/src/lookup.rs:25-41);
    `resize` then pads columns AND filters with zeros (:230-246)."""
    a = [[5, 6, 7], [1, 2, 3]]                       # 3 rows
    b = [[[5, 6, 7, 9, 9], [1, 2, 3, 9, 9]],         # table 0: 5 rows  -> trace height 5
         [[5, 5], [1, 1]]]                           # table 1: 2 rows
    blob = OT.encode_raw_lookup_trace(a, b, [0], [[1, 0]], "d")   # a_filter has 1 entry, b_filter only for table 0
    be, rows, na, nt, nb, _ = pkg.read_raw_lookup_trace(blob)
    assert (rows, na, nt, nb) == (5, 2, 2, 2)
    da, db, daf, dbf, _ = OT.decode_raw_lookup_trace(blob)
    assert daf == [0, 1, 1, 0, 0]                    # given, ones up to len(a[0]) = 3, zero padding
    assert dbf == [[1, 0, 1, 1, 1], [1, 1, 0, 0, 0]]
    assert np.array_equal(be, _be_lookup(da, db, daf, dbf, 5))


def test_lookup_malformed_is_rejected(pkg):
    a, b, af, bf = OT.synthetic_lookup_input(22, 1, 2, 4)
    obj = cbor2.loads(OT.encode_raw_lookup_trace(a, b, af, bf, "x"))
    obj["b"][1] = obj["b"][1] + obj["b"][1]          # tables of different widths
    with pytest.raises(pkg.BackendError):
        pkg.read_raw_lookup_trace(cbor2.dumps(obj))
    del obj["b"]
    with pytest.raises(pkg.BackendError):
        pkg.read_raw_lookup_trace(cbor2.dumps(obj))
    with pytest.raises(pkg.BackendError):
        pkg.read_raw_lookup_trace(OT.encode_raw_lookup_trace(a, b, af, bf, "x")[:100])
