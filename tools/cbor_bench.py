"""Host-side throughput of the CBOR reader (lsp_cbor_permutation_shape/_decode) on a cfg-5-shaped file:
6+6 columns of 2^LOG_N rows, every [u8;32] written the way serde writes it (a CBOR array of 32 small ints).
Runs without a GPU (the parser is host-only).   python tools/cbor_bench.py [LOG_N] [--write FILE]
With --write the file is kept, e.g. for `lsp_prove --permutation FILE` (BASELINE configs[4] stand-in)."""
import sys
import time

import numpy as np

sys.path.insert(0, str(__import__("pathlib").Path(__file__).resolve().parent.parent))
import __graft_entry__ as g  # noqa: E402


def head(major, n):
    if n < 24:
        return bytes([major << 5 | n])
    if n < 256:
        return bytes([major << 5 | 24, n])
    if n < 65536:
        return bytes([major << 5 | 25]) + n.to_bytes(2, "big")
    return bytes([major << 5 | 26]) + n.to_bytes(4, "big")


def column_blob(vals: np.ndarray) -> bytes:
    """vals: uint8[rows, 32] -> CBOR array(rows) of array(32) of uints, vectorised."""
    rows = vals.shape[0]
    wide = vals >= 24
    out = np.zeros((rows, 2 + 64), dtype=np.uint8)       # worst case 2 bytes per value
    lens = 2 + 32 + wide.sum(axis=1)
    out[:, 0], out[:, 1] = 0x98, 32
    pos = np.full(rows, 2)
    ar = np.arange(rows)
    for k in range(32):
        w = wide[:, k]
        out[ar[w], pos[w]] = 0x18
        pos = pos + w
        out[ar, pos] = vals[:, k]
        pos = pos + 1
    mask = np.arange(out.shape[1])[None, :] < lens[:, None]
    return head(4, rows) + out[mask].tobytes()


def main():
    argv = [a for a in sys.argv[1:]]
    write = None
    if "--write" in argv:
        k = argv.index("--write")
        write = argv[k + 1]
        del argv[k:k + 2]
    log_n = int(argv[0]) if argv else 16
    n, c = 1 << log_n, 6
    rng = np.random.default_rng(1)
    cols = [rng.integers(0, 256, size=(n, 32), dtype=np.uint8) for _ in range(c)]
    for col in cols:
        col[:, 0] &= 0x0f                               # keep values below r: no reduction needed
    perm = rng.permutation(n)
    a = b"".join(column_blob(col) for col in cols)
    b = b"".join(column_blob(col[perm]) for col in cols)
    blob = head(5, 3) + head(3, 1) + b"a" + head(4, c) + a + head(3, 1) + b"b" + head(4, c) + b + head(3, 4) + b"name" + head(3, 3) + b"mxp"
    if write:
        with open(write, "wb") as f:
            f.write(blob)
    pkg = g.load_package()
    import ctypes as C
    lib = pkg.ffi.load()
    rows_, nc_ = C.c_size_t(), C.c_uint32()
    t0 = time.perf_counter()
    assert lib.lsp_cbor_permutation_shape(blob, len(blob), C.byref(rows_), C.byref(nc_), None, 0) == 0
    t_scan = time.perf_counter() - t0
    print(f"structure pass alone: {t_scan * 1e3:.1f} ms ({len(blob) / t_scan / 1e6:.0f} MB/s)")
    t0 = time.perf_counter()
    be, rows, nc, name = pkg.read_raw_permutation_trace(blob)
    dt = time.perf_counter() - t0
    assert (rows, nc, name) == (n, c, "mxp")
    got = be.reshape(n, 2 * c, 32)
    assert all(np.array_equal(got[:, j], cols[j]) for j in range(c))
    assert all(np.array_equal(got[:, c + j], cols[j][perm]) for j in range(c))
    print(f"2^{log_n} rows x {2 * c} columns: {len(blob) / 1e6:.1f} MB of CBOR -> {be.nbytes / 1e6:.1f} MB of field bytes in "
          f"{dt * 1e3:.1f} ms ({len(blob) / dt / 1e6:.0f} MB/s of input, {n * 2 * c / dt / 1e6:.1f} M elements/s)")


if __name__ == "__main__":
    main()
