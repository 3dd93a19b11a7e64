"""Mutation fuzzing of the CBOR reader (host/cbor.cu) -- it parses files from outside the process in C++, so it must
never read out of bounds or disagree with itself: for every damaged file the serial pass and the forced parallel
pre-pass must both survive and give the same verdict and the same bytes; whenever cbor2 can still decode the file into
the struct's shape, the reader must agree with it."""
import random

import cbor2
import numpy as np
import pytest

from oracle import trace as OT


def _mutations(blob: bytes, rng: random.Random, n: int):
    out = []
    for _ in range(n):
        b = bytearray(blob)
        kind = rng.randrange(6)
        if kind == 0:                                   # truncate
            b = b[:rng.randrange(len(b))]
        elif kind == 1:                                 # flip a few bytes anywhere
            for _ in range(rng.randrange(1, 4)):
                b[rng.randrange(len(b))] = rng.randrange(256)
        elif kind == 2:                                 # damage a structural byte (array / element heads)
            heads = [i for i, x in enumerate(b) if x in (0x98, 0x9f, 0x82, 0x83, 0x84, 0xa3, 0xff, 0x18)]
            if heads:
                b[rng.choice(heads)] = rng.choice([0x98, 0x9f, 0xff, 0x1f, 0x58, 0x78, 0x18, 0x00, 0xbf, 0x9b])
        elif kind == 3:                                 # delete a slice
            i = rng.randrange(len(b))
            del b[i:i + rng.randrange(1, 40)]
        elif kind == 4:                                 # duplicate a slice
            i = rng.randrange(len(b))
            j = i + rng.randrange(1, 80)
            b[i:i] = b[i:j]
        else:                                           # huge declared length
            i = rng.randrange(len(b))
            b[i:i + 1] = bytes([0x9b]) + rng.randrange(1 << 63).to_bytes(8, "big")
        out.append(bytes(b))
    return out


def _read(pkg, blob, lookup):
    try:
        if lookup:
            buf, *meta = pkg.read_lookup_trace_once(blob)
        else:
            buf, *meta = pkg.read_permutation_trace_once(blob)
        data = buf.array.copy()
        buf.free()
        return tuple(meta), data
    except pkg.BackendError:
        return None


def _both(pkg, monkeypatch, blob, lookup):
    monkeypatch.setenv("LSP_CBOR_THREADS", "1")
    monkeypatch.delenv("LSP_CBOR_PRESCAN_MIN", raising=False)
    serial = _read(pkg, blob, lookup)
    monkeypatch.setenv("LSP_CBOR_THREADS", "4")
    monkeypatch.setenv("LSP_CBOR_PRESCAN_MIN", "0")
    parallel = _read(pkg, blob, lookup)
    assert (serial is None) == (parallel is None)
    if serial is not None:
        assert serial[0] == parallel[0] and np.array_equal(serial[1], parallel[1])
    return serial


def _expected_permutation(blob):
    """What serde would build, via cbor2 -- or None when the file is not a RawPermutationTrace."""
    try:
        obj = cbor2.loads(blob)
        a, b, name = obj["a"], obj["b"], obj["name"]
        if not isinstance(name, str) or not a or len(a) != len(b):
            return None
        cols = []
        for col in list(a) + list(b):
            vals = []
            for e in col:
                e = list(e) if isinstance(e, (list, bytes)) else None
                if e is None or len(e) != 32 or any(not isinstance(v, int) or not 0 <= v < 256 for v in e):
                    return None
                vals.append(bytes(e))
            cols.append(vals)
        rows = max(len(c) for c in cols)
        out = np.zeros((rows, len(cols), 32), dtype=np.uint8)
        for j, col in enumerate(cols):
            for i, e in enumerate(col):
                out[i, j] = np.frombuffer(e, dtype=np.uint8)
        return rows, len(a), name, out.reshape(-1)
    except Exception:                                    # noqa: BLE001
        return None


@pytest.mark.parametrize("seed", range(4))
def test_mutated_permutation_files(pkg, monkeypatch, seed):
    rng = random.Random(seed)
    a, b = OT.synthetic_permutation_input(seed, 2, 48 if seed % 2 else 33)
    blob = OT.encode_raw_permutation_trace(a, b, "fuzz")
    accepted = 0
    for bad in [blob] + _mutations(blob, rng, 150):
        got = _both(pkg, monkeypatch, bad, lookup=False)
        want = _expected_permutation(bad)
        if got is not None and want is not None and want[0] > 0:
            assert got[0] == (want[0], want[1], want[2]) and np.array_equal(got[1], want[3])
            accepted += 1
        if want is not None and want[0] > 0 and len(cbor2.dumps(cbor2.loads(bad))) == len(bad):
            assert got is not None                       # a canonical, well-formed file must be accepted
    assert accepted >= 1


@pytest.mark.parametrize("seed", range(3))
def test_mutated_lookup_files(pkg, monkeypatch, seed):
    rng = random.Random(100 + seed)
    a, b, af, bf = OT.synthetic_lookup_input(seed, 2, 2, 40, disabled_every=6)
    blob = OT.encode_raw_lookup_trace(a, b, af, bf, "fuzz")
    assert _both(pkg, monkeypatch, blob, lookup=True) is not None
    for bad in _mutations(blob, rng, 150):
        _both(pkg, monkeypatch, bad, lookup=True)          # survive, and agree with itself


# ---- the vectorised element decoder (cbor.cu: fast_elem_simd) against the byte-serial one and cbor2 ---------------
_ADVERSARIAL = [0x00, 0x17, 0x18, 0x18, 0x18, 0x19, 0x20, 0x98, 0xff]


def _element_cases(seed, n):
    """Files of one column pair whose elements are full of 0x18 (marker == payload runs), plus damaged copies: bytes of the
    element region overwritten from an alphabet that produces markers, 0x19.. heads and fake `98 20` pairs."""
    rng = random.Random(seed)
    cases = []
    for _ in range(n):
        rows = rng.choice([33, 34, 40])
        col = [[rng.choice(_ADVERSARIAL) if rng.random() < 0.8 else rng.randrange(256) for _ in range(32)] for _ in range(rows)]
        blob = bytearray(cbor2.dumps({"a": [col], "b": [col[::-1]], "name": "e"}))
        if rng.random() < 0.7:
            lo = blob.index(b"\x98\x20")
            for _ in range(rng.randrange(1, 6)):
                blob[rng.randrange(lo, len(blob) - 8)] = rng.choice(_ADVERSARIAL + [0x1a, 0x38, 0x58])
        cases.append(bytes(blob))
    return cases


def _element_digest(pkg, cases):
    import hashlib
    out = []
    for blob in cases:
        got = _read(pkg, blob, lookup=False)
        out.append(None if got is None else (got[0], hashlib.sha256(got[1].tobytes()).hexdigest()))
    return out


@pytest.mark.parametrize("seed", range(3))
def test_vector_element_decoder_equals_scalar_and_cbor2(pkg, monkeypatch, seed):
    import json
    import os
    import subprocess
    import sys
    cases = _element_cases(seed, 250)
    accepted = 0
    for blob in cases:
        got, want = _both(pkg, monkeypatch, blob, lookup=False), _expected_permutation(blob)
        if got is not None and want is not None and want[0] > 0:
            assert got[0] == (want[0], want[1], want[2]) and np.array_equal(got[1], want[3])
            accepted += 1
        if want is not None and want[0] > 0 and len(cbor2.dumps(cbor2.loads(blob))) == len(blob):
            assert got is not None
    assert accepted >= 50
    # the same files through the byte-serial decoder (a fresh process: the choice is made when the library loads)
    monkeypatch.setenv("LSP_CBOR_THREADS", "1")
    monkeypatch.delenv("LSP_CBOR_PRESCAN_MIN", raising=False)
    here = _element_digest(pkg, cases)
    code = ("import json, sys; sys.path.insert(0, %r); sys.path.insert(0, %r); import __graft_entry__ as g; import test_cbor_fuzz as t; "
            "print(json.dumps(t._element_digest(g.load_package(), t._element_cases(%d, 250))))") % (
        os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.dirname(os.path.abspath(__file__)), seed)
    env = dict(os.environ, LSP_CBOR_SCALAR="1", LSP_CBOR_THREADS="1")
    scalar = json.loads(subprocess.run([sys.executable, "-c", code], env=env, check=True, capture_output=True, text=True).stdout.splitlines()[-1])
    assert [None if d is None else [list(d[0]), d[1]] for d in here] == scalar
