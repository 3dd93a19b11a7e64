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
