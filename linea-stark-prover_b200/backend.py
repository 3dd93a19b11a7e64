"""Host-side mirror of the Plonky3 trait surface the reference selects in
`bin/src/config.rs:9-25`, implemented on the C ABI of liblsp_b200.so.

    Context                  <- construction done in bin/src/main.rs:49-66
    GpuDft.coset_lde_batch   <- TwoAdicSubgroupDft::coset_lde_batch   (Dft, config.rs:22)
    GpuMmcs.commit/open_batch<- Mmcs::commit / open_batch             (ValMmcs, config.rs:19)
    Context.permute/hash_iter<- Permutation::permute / CryptographicHasher::hash_iter
    prove                    <- p3_uni_stark::prove                   (main.rs:80-86)

Same names, argument meaning and error behaviour (the prover side of Plonky3
panics on misuse: here a `BackendError`).  Values cross as Python ints in
canonical form; the wire format underneath is Montgomery limbs (see ffi.py).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import ffi

R_MOD = 8444461749428370424248824938781546531375899335154063827935233455917409239041
_MONT_R = (1 << 256) % R_MOD
_MONT_RINV = pow(_MONT_R, -1, R_MOD)


class BackendError(RuntimeError):
    pass


def to_mont_array(values) -> np.ndarray:
    """ints (canonical) -> uint64[n,4] Montgomery limbs."""
    buf = bytearray()
    for v in values:
        buf += ((int(v) % R_MOD) * _MONT_R % R_MOD).to_bytes(32, "little")
    return np.frombuffer(bytes(buf), dtype=np.uint64).reshape(-1, 4).copy()


def from_mont_array(a: np.ndarray) -> list:
    raw = np.ascontiguousarray(a, dtype=np.uint64).tobytes()
    return [int.from_bytes(raw[i:i + 32], "little") * _MONT_RINV % R_MOD for i in range(0, len(raw), 32)]


class Context:
    """One CUDA device + stream + Poseidon2 parameters (lsp_ctx)."""

    def __init__(self, device: int = 0):
        self.lib = ffi.load()
        h = C.c_void_p()
        rc = self.lib.lsp_ctx_create(device, C.byref(h))
        if rc != 0:
            raise BackendError(f"lsp_ctx_create(device={device}) failed with {rc}: no usable CUDA device "
                               "(this backend has no CPU fallback)")
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.lsp_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc: int, what: str):
        if rc != 0:
            msg = self.lib.lsp_last_error(self.h)
            raise BackendError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")

    def sync(self):
        self.check(self.lib.lsp_ctx_sync(self.h), "lsp_ctx_sync")

    def kernel_launches(self) -> int:
        return int(self.lib.lsp_kernel_launches(self.h))

    def kernel_timing(self, enable):
        """False/0: off; True/1: every launch; 2: only the leaf-hash launches (see lsp_kernel_timing)."""
        self.check(self.lib.lsp_kernel_timing(self.h, int(enable)), "lsp_kernel_timing")

    def kernel_timing_report(self):
        import json
        buf = C.create_string_buffer(1 << 16)
        self.check(self.lib.lsp_kernel_timing_report(self.h, buf, len(buf)), "lsp_kernel_timing_report")
        return json.loads(buf.value.decode())

    def int_peak(self) -> float:
        v = C.c_double()
        self.check(self.lib.lsp_int_peak(self.h, C.byref(v)), "lsp_int_peak")
        return v.value

    INT_PEAK_FORMS = ("IMAD.WIDE.U32 (multiply only) + IADD3 + IADD3.X", "IMAD.WIDE.U32 (fused 64-bit accumulate)",
                      "IMAD.WIDE.U32.X (carry in and out)", "IMAD + IMAD.HI.U32")

    def int_peaks(self) -> dict:
        """32x32->64 multiply-accumulates per second of every instruction form (lsp_int_peaks)."""
        v = (C.c_double * len(self.INT_PEAK_FORMS))()
        self.check(self.lib.lsp_int_peaks(self.h, v), "lsp_int_peaks")
        return dict(zip(self.INT_PEAK_FORMS, [float(x) for x in v]))

    def permutation_trace(self, ab_limbs: np.ndarray, n: int, c: int, publics_limbs: np.ndarray) -> "Mat":
        """`RawPermutationTrace::get_trace` + `RawTrace::get_trace` on the device.
        ab_limbs: uint64[n*2c,4] row-major (a columns then b columns), Montgomery limbs."""
        h = C.c_void_p()
        self.check(self.lib.lsp_permutation_trace(self.h, ffi.as_u64p(ab_limbs), n, c, ffi.as_u64p(publics_limbs),
                                                  C.byref(h)), "lsp_permutation_trace")
        return Mat(self, h)

    # -- Perm::new_from_rng constants handed over (bin/src/main.rs:49) --------
    def permutation_trace_be(self, be_bytes: np.ndarray, n: int, c: int, publics_limbs: np.ndarray) -> "Mat":
        """`get_columns` + `get_trace` from raw 32-byte big-endian values (uint8[n*2c*32], row-major, a columns first)."""
        be = np.ascontiguousarray(be_bytes, dtype=np.uint8)
        assert be.size == n * 2 * c * 32
        h = C.c_void_p()
        self.check(self.lib.lsp_permutation_trace_be(self.h, be.ctypes.data, n, c, ffi.as_u64p(publics_limbs), C.byref(h)),
                   "lsp_permutation_trace_be")
        return Mat(self, h)

    def lookup_trace(self, data, n: int, n_a: int, n_tables: int, n_b: int, publics_limbs: np.ndarray) -> "Mat":
        """`RawLookupTrace::get_trace` on the device.  `data`: uint64[n*w_in,4] Montgomery limbs, or uint8 raw
        big-endian bytes (n*w_in*32), rows of  a.., b (table by table).., a_filter, b_filter[T]."""
        h = C.c_void_p()
        if data.dtype == np.uint8:
            d = np.ascontiguousarray(data)
            rc = self.lib.lsp_lookup_trace_be(self.h, d.ctypes.data, n, n_a, n_tables, n_b, ffi.as_u64p(publics_limbs), C.byref(h))
        else:
            rc = self.lib.lsp_lookup_trace(self.h, ffi.as_u64p(data), n, n_a, n_tables, n_b, ffi.as_u64p(publics_limbs), C.byref(h))
        self.check(rc, "lsp_lookup_trace")
        return Mat(self, h)

    def hconcat(self, mats) -> "Mat":
        """`RawTrace::push_traces` column concatenation."""
        arr = (C.c_void_p * len(mats))(*[m.h for m in mats])
        h = C.c_void_p()
        self.check(self.lib.lsp_mat_hconcat(self.h, arr, len(mats), C.byref(h)), "lsp_mat_hconcat")
        return Mat(self, h)

    def set_poseidon2(self, sbox_d, rounds_f, rounds_p, flat_constants, diag_m1=(1, 1, 2)):
        c = to_mont_array(flat_constants)
        d = to_mont_array(diag_m1)
        self.check(self.lib.lsp_set_poseidon2(self.h, 3, sbox_d, rounds_f, rounds_p, ffi.as_u64p(c), ffi.as_u64p(d)),
                   "lsp_set_poseidon2")

    def set_field_consts(self, generator: int, two_adic_root_2_47: int):
        """`Bls12_377Fr::GENERATOR` and `two_adic_generator(47)` (canonical ints): lsp_set_field_consts."""
        g, w = to_mont_array([generator]), to_mont_array([two_adic_root_2_47])
        self.check(self.lib.lsp_set_field_consts(self.h, ffi.as_u64p(g), ffi.as_u64p(w)), "lsp_set_field_consts")

    def set_transcript_flags(self, alpha_before_openings: bool = True, observe_opened_values: bool = False):
        """Transcript order of `TwoAdicFriPcs::open`: lsp_set_transcript_flags (pinned fork: True, False)."""
        self.check(self.lib.lsp_set_transcript_flags(self.h, int(alpha_before_openings), int(observe_opened_values)),
                   "lsp_set_transcript_flags")

    # -- parity probes ----------------------------------------------------------
    def fr_op(self, op: str, a, b=None):
        code = {"add": 0, "sub": 1, "mul": 2, "inv": 3, "halve": 4}[op]
        aa = to_mont_array(a)
        bb = to_mont_array(b) if b is not None else aa
        out = np.empty_like(aa)
        self.check(self.lib.lsp_fr_op(self.h, code, ffi.as_u64p(aa), ffi.as_u64p(bb), ffi.as_u64p(out), len(aa)),
                   "lsp_fr_op")
        return from_mont_array(out)

    def permute(self, states):
        """`Perm::permute` on a list of [Fr;3] states."""
        flat = [x for s in states for x in s]
        a = to_mont_array(flat)
        out = np.empty_like(a)
        self.check(self.lib.lsp_poseidon2_permute(self.h, ffi.as_u64p(a), ffi.as_u64p(out), len(states)),
                   "lsp_poseidon2_permute")
        o = from_mont_array(out)
        return [o[3 * i:3 * i + 3] for i in range(len(states))]

    def hash_rows(self, rows):
        """`Hash::hash_iter` applied to every row of a row-major matrix."""
        n, w = len(rows), len(rows[0]) if rows else 0
        a = to_mont_array([x for r in rows for x in r]) if w else np.zeros((0, 4), dtype=np.uint64)
        out = np.empty((n, 4), dtype=np.uint64)
        src = a if w else np.zeros((1, 4), dtype=np.uint64)
        self.check(self.lib.lsp_hash_rows(self.h, ffi.as_u64p(src), n, w, ffi.as_u64p(out)), "lsp_hash_rows")
        return from_mont_array(out)

    # -- matrices ---------------------------------------------------------------
    def upload(self, rows) -> "Mat":
        n, w = len(rows), len(rows[0])
        a = to_mont_array([x for r in rows for x in r])
        return self.upload_array(a, n, w)

    def upload_array(self, limbs: np.ndarray, n: int, w: int) -> "Mat":
        h = C.c_void_p()
        self.check(self.lib.lsp_mat_upload(self.h, ffi.as_u64p(limbs), n, w, C.byref(h)), "lsp_mat_upload")
        return Mat(self, h)


class Mat:
    """Device-resident matrix (lsp_mat).  `rows()` returns canonical ints, row-major."""

    def __init__(self, ctx: Context, h):
        self.ctx, self.h = ctx, h

    @property
    def height(self):
        return int(self.ctx.lib.lsp_mat_rows(self.h))

    @property
    def width(self):
        return int(self.ctx.lib.lsp_mat_width(self.h))

    def download_array(self, row0=0, nrows=None) -> np.ndarray:
        nrows = self.height - row0 if nrows is None else nrows
        out = np.empty((nrows * self.width, 4), dtype=np.uint64)
        self.ctx.check(self.ctx.lib.lsp_mat_download_rows(self.ctx.h, self.h, row0, nrows, ffi.as_u64p(out)),
                       "lsp_mat_download_rows")
        return out

    def rows(self, row0=0, nrows=None):
        w = self.width
        flat = from_mont_array(self.download_array(row0, nrows))
        return [flat[i:i + w] for i in range(0, len(flat), w)]

    def free(self):
        if self.h:
            self.ctx.lib.lsp_mat_free(self.ctx.h, self.h)
            self.h = None


class GpuDft:
    """`Dft` (bin/src/config.rs:22) -- only `coset_lde_batch` is on the prover path."""

    def __init__(self, ctx: Context):
        self.ctx = ctx

    def coset_lde_batch(self, mat: Mat, added_bits: int, shift: int, want_coeffs: bool = False):
        """Returns the L x W matrix in bit-reversed row order (the storage
        `.bit_reverse_rows().to_row_major_matrix()` yields)."""
        out, co = C.c_void_p(), C.c_void_p()
        s = to_mont_array([shift])
        self.ctx.check(self.ctx.lib.lsp_coset_lde_batch(self.ctx.h, mat.h, added_bits, ffi.as_u64p(s), C.byref(out),
                                                        C.byref(co) if want_coeffs else None), "lsp_coset_lde_batch")
        if want_coeffs:
            return Mat(self.ctx, out), Mat(self.ctx, co)
        return Mat(self.ctx, out)


class Tree:
    def __init__(self, ctx: Context, h, mats):
        self.ctx, self.h, self.mats = ctx, h, mats  # keep matrices alive: the tree borrows them

    @property
    def height(self):
        return int(self.ctx.lib.lsp_merkle_height(self.h))

    def layer(self, k: int):
        n = self.height >> k
        out = np.empty((n, 4), dtype=np.uint64)
        self.ctx.check(self.ctx.lib.lsp_merkle_layer(self.ctx.h, self.h, k, ffi.as_u64p(out)), "lsp_merkle_layer")
        return from_mont_array(out)

    def free(self):
        if self.h:
            self.ctx.lib.lsp_tree_free(self.ctx.h, self.h)
            self.h = None


class GpuMmcs:
    """`ValMmcs` / `ChallengeMmcs` (bin/src/config.rs:19-20)."""

    def __init__(self, ctx: Context):
        self.ctx = ctx

    def commit(self, mats):
        arr = (C.c_void_p * len(mats))(*[m.h for m in mats])
        root = np.empty((1, 4), dtype=np.uint64)
        h = C.c_void_p()
        self.ctx.check(self.ctx.lib.lsp_merkle_commit(self.ctx.h, arr, len(mats), ffi.as_u64p(root), C.byref(h)),
                       "lsp_merkle_commit")
        return from_mont_array(root)[0], Tree(self.ctx, h, list(mats))

    def open_batch(self, index: int, tree: Tree):
        widths = [m.width for m in tree.mats]
        tw = sum(widths)
        log_h = tree.height.bit_length() - 1
        rows = np.empty((tw, 4), dtype=np.uint64)
        sib = np.empty((max(log_h, 1), 4), dtype=np.uint64)
        self.ctx.check(self.ctx.lib.lsp_merkle_open_batch(self.ctx.h, tree.h, index, ffi.as_u64p(rows), ffi.as_u64p(sib)),
                       "lsp_merkle_open_batch")
        flat = from_mont_array(rows)
        opened, o = [], 0
        for w in widths:
            opened.append(flat[o:o + w])
            o += w
        return opened, from_mont_array(sib)[:log_h]

    def verify_batch(self, commit: int, log_height: int, index: int, opened_values, proof) -> bool:
        """`verify_batch(&commit, dims, index, &opened_values, &proof)`: True when the opening is consistent
        with `commit` (the reference returns `Err(RootMismatch)` otherwise)."""
        flat = [x for row in opened_values for x in row]
        rows = to_mont_array(flat)
        sib = to_mont_array(list(proof)) if log_height else np.zeros((1, 4), dtype=np.uint64)
        if len(proof) != log_height:
            raise BackendError("verify_batch: proof length != log2(height)")
        root = to_mont_array([commit])
        rc = self.ctx.lib.lsp_merkle_verify_batch(self.ctx.h, ffi.as_u64p(root), log_height, index, ffi.as_u64p(rows), len(flat),
                                                  ffi.as_u64p(sib))
        if rc == VERIFY_ROOT_MISMATCH:
            return False
        self.ctx.check(rc, "lsp_merkle_verify_batch")
        return True


# ---------------------------------------------------------------------------
# AIR config + prove (bin/src/main.rs:58-86)
# ---------------------------------------------------------------------------
class AirPermutationConfig:
    """`AirPermutationConfig` (air/src/air_permutation.rs:2-23)."""

    def __init__(self, a_columns_ids, b_columns_ids, b_inverse_id, check_id):
        self.a_columns_ids = list(a_columns_ids)
        self.b_columns_ids = list(b_columns_ids)
        self.b_inverse_id = int(b_inverse_id)
        self.check_id = int(check_id)

    def width(self):
        return len(self.a_columns_ids) + len(self.b_columns_ids) + 2


class AirLookupConfig:
    """`AirLookupConfig` (air/src/air_lookup.rs:2-39)."""

    def __init__(self, a_columns_ids, b_columns_ids, a_filter_id, b_filter_id, a_inverses_id, b_inverses_id, occurrences_id, check_id):
        self.a_columns_ids = list(a_columns_ids)
        self.b_columns_ids = [list(t) for t in b_columns_ids]
        self.a_filter_id, self.b_filter_id = int(a_filter_id), list(b_filter_id)
        self.a_inverses_id, self.b_inverses_id = int(a_inverses_id), list(b_inverses_id)
        self.occurrences_id, self.check_id = list(occurrences_id), int(check_id)

    def width(self):
        return len(self.a_columns_ids) + len(self.b_columns_ids) * (len(self.b_columns_ids[0]) + 3) + 3


class FriConfig:
    """`FriConfig` literals of bin/src/main.rs:58-64."""

    def __init__(self, log_blowup=3, log_final_poly_len=0, num_queries=33, proof_of_work_bits=0):
        self.log_blowup, self.log_final_poly_len = log_blowup, log_final_poly_len
        self.num_queries, self.proof_of_work_bits = num_queries, proof_of_work_bits

    def c_struct(self):
        return ffi.FriConfig(self.log_blowup, self.log_final_poly_len, self.num_queries, self.proof_of_work_bits)


def _c_cfgs(cfgs):
    keep = []
    arr = (ffi.PermAirCfg * len(cfgs))()
    for i, c in enumerate(cfgs):
        if len(c.a_columns_ids) != len(c.b_columns_ids):
            raise BackendError("a and b must have the same number of columns")
        a = (C.c_uint32 * len(c.a_columns_ids))(*c.a_columns_ids)
        b = (C.c_uint32 * len(c.b_columns_ids))(*c.b_columns_ids)
        keep += [a, b]
        arr[i] = ffi.PermAirCfg(len(c.a_columns_ids), a, b, c.b_inverse_id, c.check_id)
    return arr, keep


def read_raw_permutation_trace(blob: bytes, _fill=None, rows_target=None):
    """`RawPermutationTrace::read_file` (trace/src/permutation.rs:17-22) through the library's CBOR parser
    (host-only entry points: no GPU needed).  Returns (be_bytes uint8[rows*2c*32], rows, n_cols, name).
    The decoder writes every byte of the output (values and zero padding); `_fill` lets a test poison it first.
    `rows_target` >= the file's height is `push_traces`' resize to the tallest input (trace/src/lib.rs:62-79)."""
    lib = ffi.load()
    rows, nc = C.c_size_t(), C.c_uint32()
    name = C.create_string_buffer(256)
    if lib.lsp_cbor_permutation_shape(blob, len(blob), C.byref(rows), C.byref(nc), name, 256) != 0:
        raise BackendError("not a CBOR RawPermutationTrace")
    if rows_target is not None:
        rows = C.c_size_t(rows_target)
    out = np.empty(rows.value * 2 * nc.value * 32, dtype=np.uint8)
    if _fill is not None:
        out[:] = _fill
    if lib.lsp_cbor_permutation_decode(blob, len(blob), out.ctypes.data, rows.value, nc.value) != 0:
        raise BackendError("malformed CBOR RawPermutationTrace")
    return out, rows.value, nc.value, name.value.decode()


def read_raw_lookup_trace(blob: bytes, _fill=None, rows_target=None):
    """`RawLookupTrace::read_file` (trace/src/lookup.rs:20-44) through the library's CBOR parser (host only).
    Returns (be_bytes uint8[rows*(n_a + T*n_b + 1 + T)*32], rows, n_a, n_tables, n_b, name)."""
    lib = ffi.load()
    rows, na, nt, nb = C.c_size_t(), C.c_uint32(), C.c_uint32(), C.c_uint32()
    name = C.create_string_buffer(256)
    if lib.lsp_cbor_lookup_shape(blob, len(blob), C.byref(rows), C.byref(na), C.byref(nt), C.byref(nb), name, 256) != 0:
        raise BackendError("not a CBOR RawLookupTrace")
    stride = na.value + nt.value * nb.value + 1 + nt.value
    if rows_target is not None:
        rows = C.c_size_t(rows_target)
    out = np.empty(rows.value * stride * 32, dtype=np.uint8)
    if _fill is not None:
        out[:] = _fill
    if lib.lsp_cbor_lookup_decode(blob, len(blob), out.ctypes.data, rows.value, na.value, nt.value, nb.value) != 0:
        raise BackendError("malformed CBOR RawLookupTrace")
    return out, rows.value, na.value, nt.value, nb.value, name.value.decode()


class HostBuffer:
    """A buffer the library allocated (`lsp_cbor_*_read`); `.array` is a uint8 view, released by `free()` / GC."""

    def __init__(self, lib, ptr, nbytes):
        self._lib, self._ptr = lib, ptr
        self.array = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(nbytes,))

    def free(self):
        if self._ptr:
            self.array = None
            self._lib.lsp_host_free(self._ptr)
            self._ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def read_permutation_trace_once(blob: bytes):
    """`read_raw_permutation_trace` with one structure pass and a library-owned buffer (`lsp_cbor_permutation_read`).
    Returns (HostBuffer, rows, n_cols, name)."""
    lib = ffi.load()
    rows, nc, ptr = C.c_size_t(), C.c_uint32(), C.c_void_p()
    name = C.create_string_buffer(256)
    if lib.lsp_cbor_permutation_read(blob, len(blob), C.byref(rows), C.byref(nc), name, 256, C.byref(ptr)) != 0:
        raise BackendError("malformed CBOR RawPermutationTrace")
    return HostBuffer(lib, ptr, rows.value * 2 * nc.value * 32), rows.value, nc.value, name.value.decode()


def read_lookup_trace_once(blob: bytes):
    """`read_raw_lookup_trace` through `lsp_cbor_lookup_read`.  Returns (HostBuffer, rows, n_a, n_tables, n_b, name)."""
    lib = ffi.load()
    rows, na, nt, nb, ptr = C.c_size_t(), C.c_uint32(), C.c_uint32(), C.c_uint32(), C.c_void_p()
    name = C.create_string_buffer(256)
    if lib.lsp_cbor_lookup_read(blob, len(blob), C.byref(rows), C.byref(na), C.byref(nt), C.byref(nb), name, 256, C.byref(ptr)) != 0:
        raise BackendError("malformed CBOR RawLookupTrace")
    stride = na.value + nt.value * nb.value + 1 + nt.value
    return HostBuffer(lib, ptr, rows.value * stride * 32), rows.value, na.value, nt.value, nb.value, name.value.decode()


def _c_air_cfgs(cfgs):
    """Splits a `LineaAIR` config list into the (lookups, permutations) arrays of the C ABI.  The reference
    builds the list lookups-first (`RawTrace::push_traces`, trace/src/lib.rs:80-89) and the library folds the
    constraints in that order, so any other interleaving is rejected."""
    lookups = [c for c in cfgs if isinstance(c, AirLookupConfig)]
    perms = [c for c in cfgs if not isinstance(c, AirLookupConfig)]
    if list(cfgs) != lookups + perms:
        raise BackendError("AIR configs must list lookups before permutations (RawTrace::push_traces order)")
    keep = []
    larr = (ffi.LookupAirCfg * max(1, len(lookups)))()
    u32 = lambda xs: (C.c_uint32 * len(xs))(*xs)
    for i, c in enumerate(lookups):
        nb = len(c.b_columns_ids[0])
        if any(len(t) != nb for t in c.b_columns_ids):
            raise BackendError("every lookup table must have the same number of columns")
        nt = len(c.b_columns_ids)
        if not (len(c.b_filter_id) == len(c.b_inverses_id) == len(c.occurrences_id) == nt):
            raise BackendError("lookup config: per-table id lists must have one entry per table")
        bufs = [u32(c.a_columns_ids), u32([x for t in c.b_columns_ids for x in t]), u32(c.b_filter_id), u32(c.b_inverses_id),
                u32(c.occurrences_id)]
        keep += bufs
        larr[i] = ffi.LookupAirCfg(len(c.a_columns_ids), bufs[0], nt, nb, bufs[1], c.a_filter_id, bufs[2], c.a_inverses_id, bufs[3],
                                   bufs[4], c.check_id)
    parr, pkeep = _c_cfgs(perms) if perms else ((ffi.PermAirCfg * 1)(), [])
    return larr, len(lookups), parr, len(perms), keep + pkeep


STAGE_NAMES = ["commit_trace_lde", "commit_trace_merkle", "quotient", "commit_quotient", "open_reduce",
               "fri_commit_phase", "grind_query", "d2h"]


class Proof:
    """Flat proof buffer (layout in host/prover.cu) with a structured view."""

    def __init__(self, words: np.ndarray, log_n: int, width: int, log_q: int, fri: FriConfig):
        self.words, self.log_n, self.width, self.log_q, self.fri = words, log_n, width, log_q, fri

    def serialize(self) -> bytes:
        """`Proof` as bytes (lsp_proof_serialize: bincode-style field order, canonical little-endian elements)."""
        lib, cf = ffi.load(), self.fri.c_struct()
        cap = int(lib.lsp_proof_serialized_bytes(self.log_n, self.width, self.log_q, C.byref(cf)))
        buf, n = C.create_string_buffer(cap), C.c_size_t()
        rc = lib.lsp_proof_serialize(ffi.as_u64p(np.ascontiguousarray(self.words)), self.words.size, self.log_n, self.width, self.log_q,
                                     C.byref(cf), buf, cap, C.byref(n))
        if rc != 0 or n.value != cap:
            raise BackendError(f"lsp_proof_serialize failed ({rc})")
        return buf.raw

    @staticmethod
    def deserialize(blob: bytes) -> "Proof":
        """The inverse (lsp_proof_deserialize).  The per-query index slots hold the "not carried" marker: the device
        verifier fills in what it samples.  Raises `BackendError` on a malformed stream."""
        lib = ffi.load()
        log_n, width, log_q, cf, words = C.c_uint32(), C.c_uint32(), C.c_uint32(), ffi.FriConfig(), C.c_size_t()
        if lib.lsp_proof_deserialize(blob, len(blob), C.byref(log_n), C.byref(width), C.byref(log_q), C.byref(cf), None, 0, C.byref(words)) != 0:
            raise BackendError("malformed serialised proof")
        out = np.empty(words.value, dtype=np.uint64)
        if lib.lsp_proof_deserialize(blob, len(blob), C.byref(log_n), C.byref(width), C.byref(log_q), C.byref(cf), ffi.as_u64p(out),
                                     out.size, C.byref(words)) != 0:
            raise BackendError("malformed serialised proof")
        fri = FriConfig(cf.log_blowup, cf.log_final_poly_len, cf.num_queries, cf.proof_of_work_bits)
        return Proof(out, log_n.value, width.value, log_q.value, fri)

    def to_dict(self):
        """Same shape as the reference's `Proof` struct, canonical ints."""
        w, q = self.width, 1 << self.log_q
        log_l = self.log_n + self.fri.log_blowup
        rounds = self.log_n - self.fri.log_final_poly_len
        f = 1 << (self.fri.log_blowup + self.fri.log_final_poly_len)
        raw = self.words.reshape(-1, 4)
        vals = from_mont_array(raw)
        pos = 0

        def take(k):
            nonlocal pos
            out = vals[pos:pos + k]
            pos += k
            return out

        trace_commit, quot_commit = take(2)
        local, nxt = take(w), take(w)
        chunks = [[x] for x in take(q)]
        commits = take(rounds)
        final_poly = take(f)
        pow_witness = take(1)[0]
        queries, indices = [], []
        for _ in range(self.fri.num_queries):
            indices.append(int(raw[pos][0]))
            pos += 1
            ip = []
            row = take(w)
            ip.append(dict(opened_values=[row], opening_proof=take(log_l)))
            row = take(q)
            ip.append(dict(opened_values=[[x] for x in row], opening_proof=take(log_l)))
            steps = []
            for r in range(rounds):
                sib = take(1)[0]
                steps.append(dict(sibling_value=sib, opening_proof=take(log_l - 1 - r)))
            queries.append(dict(input_proof=ip, commit_phase_openings=steps))
        assert pos == len(vals)
        d = dict(commitments=dict(trace=trace_commit, quotient_chunks=quot_commit),
                 opened_values=dict(trace_local=local, trace_next=nxt, quotient_chunks=chunks),
                 opening_proof=dict(commit_phase_commits=commits, query_proofs=queries, final_poly=final_poly,
                                    pow_witness=pow_witness),
                 degree_bits=self.log_n)
        return d, indices


def prove(ctx: Context, fri: FriConfig, cfgs, trace, publics, timings=None):
    """`prove(&config, &air, &mut challenger, trace, &publics)` (bin/src/main.rs:80-86).

    `trace`: row-major list of rows (canonical ints), a uint64[n*w,4] limb array
    with `trace_shape=(n,w)` given via a tuple (array, n, w), or a device `Mat`."""
    cf = fri.c_struct()
    larr, n_l, arr, n_p, keep = _c_air_cfgs(cfgs)
    pub = to_mont_array(publics)
    assert pub.shape == (2, 4)
    tm = np.zeros(8, dtype=np.float32)
    width = sum(c.width() for c in cfgs)
    if isinstance(trace, Mat):
        n, w = trace.height, trace.width
    elif isinstance(trace, tuple):
        limbs, n, w = trace
    else:
        n, w = len(trace), len(trace[0])
        limbs = to_mont_array([x for r in trace for x in r])
    if n & (n - 1) or n == 0:
        raise BackendError(f"trace height {n} is not a power of two")
    log_n = n.bit_length() - 1
    log_q = int(ctx.lib.lsp_air_log_quotient_degree_cfg(larr, n_l, arr, n_p))
    words = int(ctx.lib.lsp_proof_words(log_n, w, log_q, C.byref(cf)))
    if words == 0:
        words = 4   # unsupported FRI parameters for this height: the call below refuses them and says why
    out = np.empty(words, dtype=np.uint64)
    if n_l == 0:      # the permutation-only entry points (the benchmarked path)
        if isinstance(trace, Mat):
            rc = ctx.lib.lsp_prove_permutation_dev(ctx.h, C.byref(cf), trace.h, arr, n_p, ffi.as_u64p(pub),
                                                   ffi.as_u64p(out), words, tm.ctypes.data_as(ffi.f32p))
        else:
            rc = ctx.lib.lsp_prove_permutation(ctx.h, C.byref(cf), ffi.as_u64p(limbs), n, w, arr, n_p,
                                               ffi.as_u64p(pub), ffi.as_u64p(out), words, tm.ctypes.data_as(ffi.f32p))
    elif isinstance(trace, Mat):
        rc = ctx.lib.lsp_prove_air_dev(ctx.h, C.byref(cf), trace.h, larr, n_l, arr, n_p, ffi.as_u64p(pub),
                                       ffi.as_u64p(out), words, tm.ctypes.data_as(ffi.f32p))
    else:
        rc = ctx.lib.lsp_prove_air(ctx.h, C.byref(cf), ffi.as_u64p(limbs), n, w, larr, n_l, arr, n_p,
                                   ffi.as_u64p(pub), ffi.as_u64p(out), words, tm.ctypes.data_as(ffi.f32p))
    ctx.check(rc, "lsp_prove_air")
    if timings is not None:
        timings.update({k: float(v) for k, v in zip(STAGE_NAMES, tm)})
    del keep, width
    return Proof(out, log_n, w, log_q, fri)


# LSP_VERIFY_* (include/lsp_b200.h): the reference verifier's rejection reasons, in the order it meets them
VERIFY_REASONS = {
    1: "InvalidProofShape",
    2: "InvalidOpeningArgument(InputError(RootMismatch)) [trace]",
    3: "InvalidOpeningArgument(InputError(RootMismatch)) [quotient chunks]",
    4: "InvalidOpeningArgument(CommitPhaseMmcsError(RootMismatch))",
    5: "InvalidOpeningArgument(FinalPolyMismatch)",
    6: "InvalidOpeningArgument(InvalidPowWitness)",
    7: "OodEvaluationMismatch",
}
VERIFY_ROOT_MISMATCH = 8


class VerificationError(Exception):
    """`p3_uni_stark::VerificationError`; `.code` is the LSP_VERIFY_* value."""

    def __init__(self, code: int):
        super().__init__(VERIFY_REASONS.get(code, f"code {code}"))
        self.code = code


def verify_code(ctx: Context, fri: FriConfig, cfgs, proof, publics, log_n=None, width=None, timing=None) -> int:
    """The raw result of `lsp_verify_air`: 0 = accepted, otherwise the LSP_VERIFY_* reason.  `proof` is a
    `Proof` or a flat uint64 word array (then `log_n` and `width` must be given)."""
    cf = fri.c_struct()
    larr, n_l, arr, n_p, keep = _c_air_cfgs(cfgs)
    pub = to_mont_array(publics)
    assert pub.shape == (2, 4)
    if isinstance(proof, Proof):
        words, log_n, width = proof.words, proof.log_n, proof.width
    else:
        words = np.ascontiguousarray(proof, dtype=np.uint64).reshape(-1)
    ms = np.zeros(1, dtype=np.float32)
    rc = ctx.lib.lsp_verify_air(ctx.h, C.byref(cf), log_n, width, larr, n_l, arr, n_p, ffi.as_u64p(pub), ffi.as_u64p(words),
                                words.size, ms.ctypes.data_as(ffi.f32p))
    del keep
    if rc < 0:
        ctx.check(rc, "lsp_verify_air")
    if timing is not None:
        timing["device_ms"] = float(ms[0])
    return int(rc)


def verify(ctx: Context, fri: FriConfig, cfgs, proof, publics, log_n=None, width=None, timing=None) -> None:
    """`verify(&config, &air, &mut challenger, &proof, &publics)` (bin/src/main.rs:88-96) on the device.
    Returns None when the proof is accepted, raises `VerificationError` otherwise."""
    rc = verify_code(ctx, fri, cfgs, proof, publics, log_n, width, timing)
    if rc:
        raise VerificationError(rc)


def quotient_permutation(ctx: Context, lde: Mat, log_n: int, log_q: int, cfgs, publics, alpha) -> Mat:
    """`quotient_values` + `split_evals`: returns the N x q matrix whose column c is chunk c."""
    arr, keep = _c_cfgs(cfgs)
    pub = to_mont_array(publics)
    al = to_mont_array([alpha])
    h = C.c_void_p()
    ctx.check(ctx.lib.lsp_quotient_permutation(ctx.h, lde.h, log_n, log_q, arr, len(cfgs), ffi.as_u64p(pub),
                                               ffi.as_u64p(al), C.byref(h)), "lsp_quotient_permutation")
    del keep
    return Mat(ctx, h)


def quotient_air(ctx: Context, lde: Mat, log_n: int, cfgs, publics, alpha) -> Mat:
    """`quotient_values` for a `LineaAIR` with lookup and permutation configs; N x q, q from the AIR's degree."""
    larr, n_l, arr, n_p, keep = _c_air_cfgs(cfgs)
    log_q = int(ctx.lib.lsp_air_log_quotient_degree_cfg(larr, n_l, arr, n_p))
    pub = to_mont_array(publics)
    al = to_mont_array([alpha])
    h = C.c_void_p()
    ctx.check(ctx.lib.lsp_quotient_air(ctx.h, lde.h, log_n, log_q, larr, n_l, arr, n_p, ffi.as_u64p(pub), ffi.as_u64p(al),
                                       C.byref(h)), "lsp_quotient_air")
    del keep
    return Mat(ctx, h)


def fri_fold(ctx: Context, vec: Mat, beta: int) -> Mat:
    """`fold_matrix(beta, m)` on a device vector (len x 1) viewed as len/2 rows of 2."""
    b = to_mont_array([beta])
    h = C.c_void_p()
    ctx.check(ctx.lib.lsp_fri_fold(ctx.h, vec.h, ffi.as_u64p(b), C.byref(h)), "lsp_fri_fold")
    return Mat(ctx, h)


def eval_at(ctx: Context, coeffs: Mat, z: int) -> list:
    """Opened values of every column at `z` from the coefficient matrix `coset_lde_batch(.., want_coeffs=True)` returned
    ("compute opened values with Lagrange interpolation", bench.log:34): lsp_eval_at."""
    out = np.empty((coeffs.width, 4), dtype=np.uint64)
    ctx.check(ctx.lib.lsp_eval_at(ctx.h, coeffs.h, ffi.as_u64p(to_mont_array([z])), ffi.as_u64p(out)), "lsp_eval_at")
    return from_mont_array(out)


def reduce_openings(ctx: Context, entries, alpha: int) -> Mat:
    """`Pcs::open`'s FRI input ("reduce rows", bench.log:35) from (lde Mat, point z, opened values ys) entries in `open`'s
    order: lsp_reduce_openings."""
    n = len(entries)
    ldes = (C.c_void_p * n)(*[m.h for m, _, _ in entries])
    pts = to_mont_array([z for _, z, _ in entries])
    ys = [to_mont_array(y) for _, _, y in entries]
    yptr = (ffi.u64p * n)(*[ffi.as_u64p(y) for y in ys])
    h = C.c_void_p()
    ctx.check(ctx.lib.lsp_reduce_openings(ctx.h, ldes, ffi.as_u64p(pts), yptr, n, ffi.as_u64p(to_mont_array([alpha])), C.byref(h)),
              "lsp_reduce_openings")
    return Mat(ctx, h)


# ---------------------------------------------------------------------------
# multi-GPU: one proof sharded over ranks
# ---------------------------------------------------------------------------
class Comm:
    """`lsp_comm`: either NCCL over one process per GPU, or all ranks emulated on one device."""

    def __init__(self, ctx: Context, h, world: int, rank: int):
        self.ctx, self.h, self.world, self.rank = ctx, h, world, rank

    @staticmethod
    def local(ctx: Context, world: int) -> "Comm":
        h = C.c_void_p()
        ctx.check(ctx.lib.lsp_comm_init_local(ctx.h, world, C.byref(h)), "lsp_comm_init_local")
        return Comm(ctx, h, world, 0)

    @staticmethod
    def nccl(ctx: Context, rank: int, world: int, unique_id: bytes) -> "Comm":
        assert len(unique_id) == 128
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        h = C.c_void_p()
        ctx.check(ctx.lib.lsp_comm_init_nccl(ctx.h, rank, world, buf, C.byref(h)), "lsp_comm_init_nccl")
        return Comm(ctx, h, world, rank)

    @staticmethod
    def unique_id() -> bytes:
        buf = (C.c_uint8 * 128)()
        rc = ffi.load().lsp_nccl_unique_id(buf)
        if rc != 0:
            raise BackendError(f"lsp_nccl_unique_id failed ({rc}): libnccl.so.2 not loadable")
        return bytes(buf)

    def close(self):
        if self.h:
            self.ctx.lib.lsp_comm_destroy(self.h)
            self.h = None


def shard_plan(log_n: int, log_blowup: int, world: int, rank: int) -> dict:
    """Which part of the committed matrices a rank owns: storage rows [row0, row0+rows) of the
    bit-reversed LDE = `cosets` (exponents c of shift*w_L^c) of the evaluation domain."""
    big = 1 << (log_n + log_blowup)
    if world < 1 or world & (world - 1):
        raise BackendError(f"{world} ranks: the rank count must be a power of two")
    rows = big // world
    rev = lambda x, bits: int(format(x, f"0{bits}b")[::-1], 2) if bits else 0
    if world <= (1 << log_blowup):      # whole cosets
        blocks = (1 << log_blowup) // world
        return dict(row0=rank * rows, rows=rows, blocks=list(range(rank * blocks, (rank + 1) * blocks)),
                    cosets=[rev(b, log_blowup) for b in range(rank * blocks, (rank + 1) * blocks)], fraction=(0, 1))
    # more ranks than cosets: the fraction 1/S of one coset -- the trace rows k = k0 (mod S) of it, i.e. the sub-coset
    # sigma * H_{N/S} with sigma = shift * w_L^c * w_N^k0 (host/sharded.cu, csrc/ntt.cu coset_evaluate_subblock)
    log_s = (world >> log_blowup).bit_length() - 1
    if log_n - log_s < 3:
        raise BackendError(f"{world} ranks are too many for 2^{log_n} x 2^{log_blowup} rows")
    block, sub = rank >> log_s, rank & ((1 << log_s) - 1)
    return dict(row0=rank * rows, rows=rows, blocks=[block], cosets=[rev(block, log_blowup)], fraction=(rev(sub, log_s), 1 << log_s))


def prove_sharded(comm: Comm, fri: FriConfig, cfgs, trace, publics, timings=None):
    """`prove` with the LDE / Merkle / FRI work sharded over comm's ranks.  Same proof as `prove`."""
    ctx = comm.ctx
    cf = fri.c_struct()
    larr, n_l, arr, n_p, keep = _c_air_cfgs(cfgs)
    pub = to_mont_array(publics)
    tm = np.zeros(8, dtype=np.float32)
    if isinstance(trace, Mat):
        n, w = trace.height, trace.width
    elif isinstance(trace, tuple):
        limbs, n, w = trace
    else:
        n, w = len(trace), len(trace[0])
        limbs = to_mont_array([x for r in trace for x in r])
    if n & (n - 1) or n == 0:
        raise BackendError(f"trace height {n} is not a power of two")
    log_n = n.bit_length() - 1
    log_q = int(ctx.lib.lsp_air_log_quotient_degree_cfg(larr, n_l, arr, n_p))
    words = int(ctx.lib.lsp_proof_words(log_n, w, log_q, C.byref(cf)))
    if words == 0:
        words = 4   # unsupported FRI parameters for this height: the call below refuses them and says why
    out = np.empty(words, dtype=np.uint64)
    if isinstance(trace, Mat):
        rc = ctx.lib.lsp_prove_air_sharded_dev(comm.h, C.byref(cf), trace.h, larr, n_l, arr, n_p, ffi.as_u64p(pub),
                                               ffi.as_u64p(out), words, tm.ctypes.data_as(ffi.f32p))
    else:
        rc = ctx.lib.lsp_prove_air_sharded(comm.h, C.byref(cf), ffi.as_u64p(limbs), n, w, larr, n_l, arr, n_p,
                                           ffi.as_u64p(pub), ffi.as_u64p(out), words, tm.ctypes.data_as(ffi.f32p))
    ctx.check(rc, "lsp_prove_air_sharded")
    if timings is not None:
        timings.update({k: float(v) for k, v in zip(STAGE_NAMES, tm)})
    del keep
    return Proof(out, log_n, w, log_q, fri)
