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
    be, rows, nc, _ = pkg.read_raw_permutation_trace(OT.encode_raw_permutation_trace(a, b, "r"), _fill=0xEE)
    assert (rows, nc) == (4, 2)
    assert np.array_equal(be, _be([a[0], a[1] + [0, 0]], [b[0] + [0, 0, 0], b[1] + [0]], 4))


def test_fast_numpy_encoder_is_the_oracle_encoder(pkg):
    """The vectorised encoder the full-size tests use writes the same bytes as the oracle's cbor2 encoder."""
    from tests.proofs import fast_cbor_permutation_trace
    for n, c, seed in ((1, 1, 1), (23, 2, 2), (24, 3, 3), (300, 1, 4), (70000, 1, 5)):
        a, b = OT.synthetic_permutation_input(seed, c, n) if n < 1000 else ([[(i * 2654435761) % F.R_MOD for i in range(n)]], [[i for i in range(n)]])
        assert fast_cbor_permutation_trace(_be(a, b, n), n, c, "nm") == OT.encode_raw_permutation_trace(a, b, "nm")


def test_decode_to_a_taller_common_height_pads_with_zero_rows(pkg):
    """`push_traces` (trace/src/lib.rs:62-79) resizes every sub-trace to the tallest input before its witness is built:
    `_decode` accepts rows >= the file's height, `_read_rows` allocates max(min_rows, height); a shorter target is an error."""
    import ctypes as C
    a, b = [[1, 2, 3], [4, 5]], [[3, 2, 1], [5, 4, 0]]
    blob = OT.encode_raw_permutation_trace(a, b, "short")
    be, rows, nc, _ = pkg.read_raw_permutation_trace(blob, _fill=0xEE, rows_target=8)
    assert (rows, nc) == (8, 2)
    assert np.array_equal(be, _be([a[0] + [0] * 5, a[1] + [0] * 6], [b[0] + [0] * 5, b[1] + [0] * 5], 8))
    with pytest.raises(pkg.BackendError):
        pkg.read_raw_permutation_trace(blob, rows_target=2)
    lib = pkg.ffi.load()
    r, n, ptr = C.c_size_t(), C.c_uint32(), C.c_void_p()
    assert lib.lsp_cbor_permutation_read_rows(blob, len(blob), 8, C.byref(r), C.byref(n), None, 0, C.byref(ptr)) == 0
    got = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(8 * 4 * 32,)).copy()
    lib.lsp_host_free(ptr)
    assert (r.value, n.value) == (8, 2) and np.array_equal(got, be)
    # lookups: the filters of the padded rows are ZERO (resize, trace/src/lookup.rs:230-246), not the default one
    lk = OT.synthetic_lookup_input(5, 2, 1, 4)
    lblob = OT.encode_raw_lookup_trace(*lk, "lk")
    lbe, lrows, na, nt, nb, _ = pkg.read_raw_lookup_trace(lblob, _fill=0xEE, rows_target=8)
    plain, *_ = pkg.read_raw_lookup_trace(lblob)
    stride = na + nt * nb + 1 + nt
    assert lrows == 8 and np.array_equal(lbe[:4 * stride * 32], plain) and not lbe[4 * stride * 32:].any()


def test_deeply_nested_unknown_value_is_rejected_not_a_stack_overflow(pkg):
    """An unknown key whose value nests deeper than ciborium's recursion limit (256): a malformed file, not a crash."""
    e = lambda x: int(x).to_bytes(32, "big")
    good = cbor2.dumps({"a": [[e(1)]], "b": [[e(1)]], "name": "x"})
    for opener, closer in ((b"\x81", b""), (b"\xc0", b""), (b"\x9f", b"\xff")):
        for depth, ok in ((200, True), (100_000, False)):
            junk = opener * depth + b"\x00" + closer * depth
            blob = b"\xa4" + good[1:] + cbor2.dumps("zz") + junk
            if ok:
                assert pkg.read_raw_permutation_trace(blob)[1] == 1
            else:
                with pytest.raises(pkg.BackendError):
                    pkg.read_raw_permutation_trace(blob)


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
    be, rows, na, nt, nb, _ = pkg.read_raw_lookup_trace(blob, _fill=0xEE)
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


# ---- the parallel structure pre-pass (host/cbor.cu: prescan) must be invisible ------------------------------
def _both_ways(monkeypatch, read, blob):
    """Reads `blob` with the serial pass and with the parallel pre-pass forced on (tiny chunks, 5 threads)."""
    monkeypatch.setenv("LSP_CBOR_THREADS", "1")
    monkeypatch.delenv("LSP_CBOR_PRESCAN_MIN", raising=False)
    try:
        serial = read(blob)
    except Exception as e:                                   # noqa: BLE001
        serial = type(e)
    monkeypatch.setenv("LSP_CBOR_THREADS", "5")
    monkeypatch.setenv("LSP_CBOR_PRESCAN_MIN", "0")
    try:
        parallel = read(blob)
    except Exception as e:                                   # noqa: BLE001
        parallel = type(e)
    return serial, parallel


def _same(x, y):
    if isinstance(x, type) or isinstance(y, type):
        return x is y
    return all(np.array_equal(a, b) if isinstance(a, np.ndarray) else a == b for a, b in zip(x, y))


@pytest.mark.parametrize("rows,c,ragged", [(3, 1, False), (32, 2, False), (33, 2, True), (200, 3, True), (1000, 6, False)])
def test_parallel_prescan_equals_serial_permutation(pkg, monkeypatch, rows, c, ragged):
    """32-row columns have the element marker `98 20` as their own head: the pre-pass must notice and step aside."""
    a, b = OT.synthetic_permutation_input(40 + rows, c, rows)
    a, b = [list(col) for col in a], [list(col) for col in b]
    if ragged:
        a[0] = a[0][:rows // 2]
        b[-1] = b[-1][:1]
    a[0][0] = 0x9820 << 8 | 0x18                              # marker-looking bytes inside an element
    blob = OT.encode_raw_permutation_trace(a, b, "pؘ q")  # ... and inside the name (d8 98 20)
    poisoned = lambda x: pkg.read_raw_permutation_trace(x, _fill=0xCD)      # every output byte must be written
    serial, parallel = _both_ways(monkeypatch, poisoned, blob)
    assert not isinstance(serial, type) and _same(serial, parallel)
    assert serial[1:] == (rows, c, "pؘ q")
    pad = lambda col: list(col) + [0] * (rows - len(col))
    assert np.array_equal(serial[0], _be([pad(x) for x in a], [pad(x) for x in b], rows))


def test_parallel_prescan_equals_serial_lookup(pkg, monkeypatch):
    a, b, af, bf = OT.synthetic_lookup_input(61, 2, 3, 300, disabled_every=7)
    blob = OT.encode_raw_lookup_trace(a, b, af[:100], bf[:2], "lk")      # short a_filter, one table's filter missing
    serial, parallel = _both_ways(monkeypatch, lambda x: pkg.read_raw_lookup_trace(x, _fill=0xCD), blob)
    assert not isinstance(serial, type) and _same(serial, parallel)
    da, db, daf, dbf, _ = OT.decode_raw_lookup_trace(blob)
    assert np.array_equal(serial[0], _be_lookup(da, db, daf, dbf, 300))


@pytest.mark.parametrize("damage", ["truncate", "flip_head", "byte_strings", "indefinite", "extra_key", "duplicate_key"])
def test_parallel_prescan_on_irregular_and_malformed_files(pkg, monkeypatch, damage):
    """Whatever the pre-pass makes of a file, the verdict and the bytes are the serial pass's."""
    a, b = OT.synthetic_permutation_input(77, 2, 150)
    blob = bytearray(OT.encode_raw_permutation_trace(a, b, "z"))
    if damage == "truncate":
        blob = blob[:len(blob) * 2 // 3]
    elif damage == "flip_head":
        k = blob.index(b"\x98\x20", len(blob) // 2)
        blob[k + 1] = 0x1f                                    # array(31) where an element should be
    elif damage == "byte_strings":
        e = lambda x: int(x).to_bytes(32, "big")
        blob = cbor2.dumps({"a": [[e(x) for x in col] for col in a], "b": [[e(x) for x in col] for col in b], "name": "z"})
    elif damage == "indefinite":
        k = blob.index(b"\x98\x20", len(blob) // 3)
        elem_end = k + 2
        for _ in range(32):
            elem_end += 2 if blob[elem_end] == 0x18 else 1
        blob = blob[:k] + b"\x9f" + blob[k + 2:elem_end] + b"\xff" + blob[elem_end:]
    elif damage == "extra_key":
        obj = cbor2.loads(bytes(blob))
        obj["zz"] = [[list(range(32))] * 40]                  # element-shaped data under a key the struct does not have
        blob = cbor2.dumps(obj)
    elif damage == "duplicate_key":
        tail = bytes(blob[blob.index(b"aa"):blob.index(b"ab")])       # key "a" .. up to key "b": the whole `a` entry
        blob = bytes([blob[0] + 1]) + bytes(blob[1:]) + tail
    serial, parallel = _both_ways(monkeypatch, pkg.read_raw_permutation_trace, bytes(blob))
    assert _same(serial, parallel)
    assert isinstance(serial, type) == (damage in ("truncate", "flip_head", "duplicate_key"))


def test_single_pass_read_equals_shape_plus_decode(pkg, monkeypatch):
    """`lsp_cbor_*_read` (library-allocated buffer, one structure pass) == `_shape` + `_decode`, serial and parallel."""
    a, b = OT.synthetic_permutation_input(5, 3, 120)
    a[1] = a[1][:70]
    pblob = OT.encode_raw_permutation_trace(a, b, "one")
    la, lb, laf, lbf = OT.synthetic_lookup_input(6, 2, 2, 90, disabled_every=5)
    lblob = OT.encode_raw_lookup_trace(la, lb, laf[:10], lbf[:1], "two")
    for threads, prescan in (("1", None), ("6", "0")):
        monkeypatch.setenv("LSP_CBOR_THREADS", threads)
        if prescan is None:
            monkeypatch.delenv("LSP_CBOR_PRESCAN_MIN", raising=False)
        else:
            monkeypatch.setenv("LSP_CBOR_PRESCAN_MIN", prescan)
        want = pkg.read_raw_permutation_trace(pblob, _fill=0x5A)
        buf, *meta = pkg.read_permutation_trace_once(pblob)
        assert tuple(meta) == want[1:] and np.array_equal(buf.array, want[0])
        buf.free()
        want = pkg.read_raw_lookup_trace(lblob, _fill=0x5A)
        buf, *meta = pkg.read_lookup_trace_once(lblob)
        assert tuple(meta) == want[1:] and np.array_equal(buf.array, want[0])
        buf.free()
    for bad in (pblob[:200], b"\x00", lblob[:-3]):
        with pytest.raises(pkg.BackendError):
            pkg.read_permutation_trace_once(bad)
    with pytest.raises(pkg.BackendError):
        pkg.read_lookup_trace_once(lblob[:300])


def test_pinned_output_buffers_fall_back_to_ordinary_memory_without_a_device(pkg):
    """`lsp_host_pinned(1)` asks for page-locked `_read` buffers; where nothing can be pinned (this CPU-only container) the
    reader must still work on an ordinary buffer, and blocks handed back must be reusable."""
    lib = pkg.ffi.load()
    a, b = OT.synthetic_permutation_input(9, 2, 40)
    blob = OT.encode_raw_permutation_trace(a, b, "pin")
    want = pkg.read_raw_permutation_trace(blob)
    assert lib.lsp_host_pinned(1) == 0
    try:
        for _ in range(3):                                   # alloc, release into the pool, alloc again
            buf, rows, nc, name = pkg.read_permutation_trace_once(blob)
            assert (rows, nc, name) == want[1:] and np.array_equal(buf.array, want[0])
            buf.free()
    finally:
        assert lib.lsp_host_pinned(0) == 0
    buf, *_ = pkg.read_permutation_trace_once(blob)
    assert np.array_equal(buf.array, want[0])
    buf.free()
