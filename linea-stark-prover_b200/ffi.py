"""ctypes binding of liblsp_b200.so (the C ABI in include/lsp_b200.h).

This is the only way Python reaches the product: every call goes through the
same extern "C" entry points a Rust shim would bind.  There is no fallback --
a missing library or a missing CUDA device raises.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

PKG_DIR = Path(__file__).resolve().parent
# LSP_B200_LIB selects another build of the same library (tuning experiments); the default is the in-tree one
LIB_PATH = Path(os.environ.get("LSP_B200_LIB") or PKG_DIR / "liblsp_b200.so")

u64p = C.POINTER(C.c_uint64)
f32p = C.POINTER(C.c_float)
vp = C.c_void_p


class PermAirCfg(C.Structure):
    """`lsp_perm_air_cfg` == reference `AirPermutationConfig` (air/src/air_permutation.rs:2-7)."""
    _fields_ = [("n_cols", C.c_uint32), ("a_ids", C.POINTER(C.c_uint32)), ("b_ids", C.POINTER(C.c_uint32)),
                ("b_inverse_id", C.c_uint32), ("check_id", C.c_uint32)]


class LookupAirCfg(C.Structure):
    """`lsp_lookup_air_cfg` == reference `AirLookupConfig` (air/src/air_lookup.rs:2-11)."""
    _fields_ = [("n_a_cols", C.c_uint32), ("a_ids", C.POINTER(C.c_uint32)), ("n_tables", C.c_uint32), ("n_b_cols", C.c_uint32),
                ("b_ids", C.POINTER(C.c_uint32)), ("a_filter_id", C.c_uint32), ("b_filter_ids", C.POINTER(C.c_uint32)),
                ("a_inverses_id", C.c_uint32), ("b_inverses_ids", C.POINTER(C.c_uint32)),
                ("occurrences_ids", C.POINTER(C.c_uint32)), ("check_id", C.c_uint32)]


class FriConfig(C.Structure):
    """`lsp_fri_config` == reference `FriConfig` literals (bin/src/main.rs:58-64)."""
    _fields_ = [("log_blowup", C.c_uint32), ("log_final_poly_len", C.c_uint32), ("num_queries", C.c_uint32),
                ("proof_of_work_bits", C.c_uint32)]


# name -> (restype, argtypes); kept in one table so tests can check that every
# symbol the header declares is exported.
SIGNATURES = {
    "lsp_abi_version": (C.c_int, []),
    "lsp_ctx_create": (C.c_int, [C.c_int, C.POINTER(vp)]),
    "lsp_ctx_destroy": (None, [vp]),
    "lsp_last_error": (C.c_char_p, [vp]),
    "lsp_ctx_sync": (C.c_int, [vp]),
    "lsp_kernel_launches": (C.c_uint64, [vp]),
    "lsp_kernel_timing": (C.c_int, [vp, C.c_int]),
    "lsp_kernel_timing_report": (C.c_int, [vp, C.c_char_p, C.c_size_t]),
    "lsp_int_peak": (C.c_int, [vp, C.POINTER(C.c_double)]),
    "lsp_int_peaks": (C.c_int, [vp, C.POINTER(C.c_double)]),
    "lsp_permutation_trace": (C.c_int, [vp, u64p, C.c_size_t, C.c_uint32, u64p, C.POINTER(vp)]),
    "lsp_set_poseidon2": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_int, u64p, u64p]),
    "lsp_set_field_consts": (C.c_int, [vp, u64p, u64p]),
    "lsp_set_transcript_flags": (C.c_int, [vp, C.c_int, C.c_int]),
    "lsp_fr_op": (C.c_int, [vp, C.c_int, u64p, u64p, u64p, C.c_size_t]),
    "lsp_poseidon2_permute": (C.c_int, [vp, u64p, u64p, C.c_size_t]),
    "lsp_hash_rows": (C.c_int, [vp, u64p, C.c_size_t, C.c_size_t, u64p]),
    "lsp_mat_upload": (C.c_int, [vp, u64p, C.c_size_t, C.c_size_t, C.POINTER(vp)]),
    "lsp_mat_download": (C.c_int, [vp, vp, u64p]),
    "lsp_mat_download_rows": (C.c_int, [vp, vp, C.c_size_t, C.c_size_t, u64p]),
    "lsp_mat_rows": (C.c_size_t, [vp]),
    "lsp_mat_width": (C.c_size_t, [vp]),
    "lsp_mat_free": (None, [vp, vp]),
    "lsp_coset_lde_batch": (C.c_int, [vp, vp, C.c_int, u64p, C.POINTER(vp), C.POINTER(vp)]),
    "lsp_merkle_commit": (C.c_int, [vp, C.POINTER(vp), C.c_int, u64p, C.POINTER(vp)]),
    "lsp_merkle_open_batch": (C.c_int, [vp, vp, C.c_size_t, u64p, u64p]),
    "lsp_merkle_layer": (C.c_int, [vp, vp, C.c_int, u64p]),
    "lsp_merkle_height": (C.c_size_t, [vp]),
    "lsp_tree_free": (None, [vp, vp]),
    "lsp_quotient_permutation": (C.c_int, [vp, vp, C.c_int, C.c_int, C.POINTER(PermAirCfg), C.c_int, u64p, u64p,
                                           C.POINTER(vp)]),
    "lsp_eval_at": (C.c_int, [vp, vp, u64p, u64p]),
    "lsp_reduce_openings": (C.c_int, [vp, C.POINTER(vp), u64p, C.POINTER(u64p), C.c_int, u64p, C.POINTER(vp)]),
    "lsp_fri_fold": (C.c_int, [vp, vp, u64p, C.POINTER(vp)]),
    "lsp_proof_words": (C.c_size_t, [C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(FriConfig)]),
    "lsp_prove_permutation": (C.c_int, [vp, C.POINTER(FriConfig), u64p, C.c_size_t, C.c_size_t,
                                        C.POINTER(PermAirCfg), C.c_int, u64p, u64p, C.c_size_t, f32p]),
    "lsp_prove_permutation_dev": (C.c_int, [vp, C.POINTER(FriConfig), vp, C.POINTER(PermAirCfg), C.c_int, u64p,
                                            u64p, C.c_size_t, f32p]),
    "lsp_proof_serialized_bytes": (C.c_size_t, [C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(FriConfig)]),
    "lsp_proof_serialize": (C.c_int, [u64p, C.c_size_t, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(FriConfig), C.c_char_p, C.c_size_t,
                                      C.POINTER(C.c_size_t)]),
    "lsp_proof_deserialize": (C.c_int, [C.c_char_p, C.c_size_t, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32),
                                        C.POINTER(FriConfig), u64p, C.c_size_t, C.POINTER(C.c_size_t)]),
    "lsp_nccl_unique_id": (C.c_int, [C.POINTER(C.c_uint8)]),
    "lsp_comm_init_nccl": (C.c_int, [vp, C.c_int, C.c_int, C.POINTER(C.c_uint8), C.POINTER(vp)]),
    "lsp_comm_init_local": (C.c_int, [vp, C.c_int, C.POINTER(vp)]),
    "lsp_comm_destroy": (None, [vp]),
    "lsp_prove_air_sharded": (C.c_int, [vp, C.POINTER(FriConfig), u64p, C.c_size_t, C.c_size_t, C.POINTER(LookupAirCfg), C.c_int,
                                        C.POINTER(PermAirCfg), C.c_int, u64p, u64p, C.c_size_t, f32p]),
    "lsp_prove_air_sharded_dev": (C.c_int, [vp, C.POINTER(FriConfig), vp, C.POINTER(LookupAirCfg), C.c_int, C.POINTER(PermAirCfg),
                                            C.c_int, u64p, u64p, C.c_size_t, f32p]),
    "lsp_air_log_quotient_degree": (C.c_int, [C.c_int, C.c_int]),
    "lsp_air_log_quotient_degree_cfg": (C.c_int, [C.POINTER(LookupAirCfg), C.c_int, C.POINTER(PermAirCfg), C.c_int]),
    "lsp_cbor_permutation_shape": (C.c_int, [C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_uint32), C.c_char_p, C.c_size_t]),
    "lsp_cbor_permutation_decode": (C.c_int, [C.c_char_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_uint32]),
    "lsp_cbor_lookup_shape": (C.c_int, [C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32),
                                        C.POINTER(C.c_uint32), C.c_char_p, C.c_size_t]),
    "lsp_cbor_lookup_decode": (C.c_int, [C.c_char_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint32, C.c_uint32]),
    "lsp_cbor_permutation_read": (C.c_int, [C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_uint32), C.c_char_p, C.c_size_t,
                                            C.POINTER(C.c_void_p)]),
    "lsp_cbor_lookup_read": (C.c_int, [C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32),
                                       C.POINTER(C.c_uint32), C.c_char_p, C.c_size_t, C.POINTER(C.c_void_p)]),
    "lsp_cbor_permutation_read_rows": (C.c_int, [C.c_char_p, C.c_size_t, C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_uint32), C.c_char_p,
                                                 C.c_size_t, C.POINTER(C.c_void_p)]),
    "lsp_cbor_lookup_read_rows": (C.c_int, [C.c_char_p, C.c_size_t, C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_uint32),
                                            C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.c_char_p, C.c_size_t, C.POINTER(C.c_void_p)]),
    "lsp_host_free": (None, [C.c_void_p]),
    "lsp_host_pinned": (C.c_int, [C.c_int]),
    "lsp_lookup_trace": (C.c_int, [vp, u64p, C.c_size_t, C.c_uint32, C.c_uint32, C.c_uint32, u64p, C.POINTER(vp)]),
    "lsp_lookup_trace_be": (C.c_int, [vp, C.c_void_p, C.c_size_t, C.c_uint32, C.c_uint32, C.c_uint32, u64p, C.POINTER(vp)]),
    "lsp_mat_hconcat": (C.c_int, [vp, C.POINTER(vp), C.c_int, C.POINTER(vp)]),
    "lsp_permutation_trace_be": (C.c_int, [vp, C.c_void_p, C.c_size_t, C.c_uint32, u64p, C.POINTER(vp)]),
    "lsp_quotient_air": (C.c_int, [vp, vp, C.c_int, C.c_int, C.POINTER(LookupAirCfg), C.c_int, C.POINTER(PermAirCfg), C.c_int,
                                   u64p, u64p, C.POINTER(vp)]),
    "lsp_prove_air": (C.c_int, [vp, C.POINTER(FriConfig), u64p, C.c_size_t, C.c_size_t, C.POINTER(LookupAirCfg), C.c_int,
                                C.POINTER(PermAirCfg), C.c_int, u64p, u64p, C.c_size_t, f32p]),
    "lsp_prove_air_dev": (C.c_int, [vp, C.POINTER(FriConfig), vp, C.POINTER(LookupAirCfg), C.c_int, C.POINTER(PermAirCfg), C.c_int,
                                    u64p, u64p, C.c_size_t, f32p]),
    "lsp_verify_air": (C.c_int, [vp, C.POINTER(FriConfig), C.c_uint32, C.c_size_t, C.POINTER(LookupAirCfg), C.c_int,
                                 C.POINTER(PermAirCfg), C.c_int, u64p, u64p, C.c_size_t, f32p]),
    "lsp_verify_permutation": (C.c_int, [vp, C.POINTER(FriConfig), C.c_uint32, C.c_size_t, C.POINTER(PermAirCfg), C.c_int,
                                         u64p, u64p, C.c_size_t, f32p]),
    "lsp_merkle_verify_batch": (C.c_int, [vp, u64p, C.c_uint32, C.c_size_t, u64p, C.c_size_t, u64p]),
    "lsp_prove_permutation_sharded": (C.c_int, [vp, C.POINTER(FriConfig), u64p, C.c_size_t, C.c_size_t,
                                                C.POINTER(PermAirCfg), C.c_int, u64p, u64p, C.c_size_t, f32p]),
    "lsp_prove_permutation_sharded_dev": (C.c_int, [vp, C.POINTER(FriConfig), vp, C.POINTER(PermAirCfg), C.c_int, u64p, u64p,
                                                    C.c_size_t, f32p]),
}

_lib = None


def library_path() -> Path:
    return LIB_PATH


def load() -> C.CDLL:
    """dlopen the in-tree library and type every entry point.  Raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback)")
    lib = C.CDLL(str(LIB_PATH), mode=os.RTLD_NOW | os.RTLD_LOCAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the export is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def as_u64p(a: np.ndarray):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(u64p)
