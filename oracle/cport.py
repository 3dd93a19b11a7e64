"""ctypes binding of oracle/c/liblsp_oracle.so -- the multi-threaded C twin of the
Python oracle (TEST INFRASTRUCTURE ONLY; see oracle/__init__.py)."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

from .field import to_mont_limbs, from_mont_limbs

LIB = Path(__file__).resolve().parent / "c" / "liblsp_oracle.so"
u64p = C.POINTER(C.c_uint64)
f64p = C.POINTER(C.c_double)


class AirCfg(C.Structure):
    _fields_ = [("n_cols", C.c_uint32), ("a_ids", C.POINTER(C.c_uint32)), ("b_ids", C.POINTER(C.c_uint32)),
                ("b_inverse_id", C.c_uint32), ("check_id", C.c_uint32)]


class FriCfg(C.Structure):
    _fields_ = [("log_blowup", C.c_uint32), ("log_final_poly_len", C.c_uint32), ("num_queries", C.c_uint32),
                ("proof_of_work_bits", C.c_uint32)]


_lib = None


def load():
    global _lib
    if _lib is None:
        if not LIB.exists():
            raise RuntimeError(f"{LIB} missing: run `make -C oracle/c` (or __graft_entry__.build())")
        lib = C.CDLL(str(LIB))
        lib.lsp_oracle_set_poseidon2.argtypes = [C.c_int, C.c_int, C.c_int, u64p, u64p]
        lib.lsp_oracle_permute.argtypes = [u64p, u64p, C.c_size_t]
        lib.lsp_oracle_fr_mul.argtypes = [u64p, u64p, u64p, C.c_size_t]
        lib.lsp_oracle_proof_words.restype = C.c_size_t
        lib.lsp_oracle_proof_words.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(FriCfg)]
        lib.lsp_oracle_prove.argtypes = [C.POINTER(FriCfg), u64p, C.c_size_t, C.c_size_t, C.POINTER(AirCfg), C.c_int,
                                         u64p, u64p, C.c_size_t, f64p]
        lib.lsp_oracle_verify.argtypes = [C.POINTER(FriCfg), C.c_uint32, C.c_size_t, C.POINTER(AirCfg), C.c_int, u64p,
                                          u64p, C.c_size_t]
        lib.lsp_oracle_gen_trace.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, u64p, u64p]
        lib.lsp_oracle_permutation_trace.argtypes = [u64p, C.c_size_t, C.c_uint32, u64p, u64p]
        lib.lsp_oracle_set_field_consts.argtypes = [u64p, u64p]
        lib.lsp_oracle_set_transcript_flags.argtypes = [C.c_int, C.c_int]
        lib.lsp_oracle_set_lookups.argtypes = [C.POINTER(C.c_uint32), C.c_int]
        _lib = lib
    return _lib


def _arr(vals):
    return np.array([to_mont_limbs(v) for v in vals], dtype=np.uint64).reshape(-1, 4)


def _p(a):
    return a.ctypes.data_as(u64p)


def set_poseidon2(p):
    c = _arr(p.flat_constants())
    d = _arr(p.internal_diag_m1)
    assert load().lsp_oracle_set_poseidon2(p.sbox_d, p.rounds_f, p.rounds_p, _p(c), _p(d)) == 0


def permute(states):
    a = _arr([x for s in states for x in s])
    out = np.empty_like(a)
    load().lsp_oracle_permute(_p(a), _p(out), len(states))
    flat = [from_mont_limbs(r) for r in out]
    return [flat[3 * i:3 * i + 3] for i in range(len(states))]


def fr_mul(a, b):
    aa, bb = _arr(a), _arr(b)
    out = np.empty_like(aa)
    load().lsp_oracle_fr_mul(_p(aa), _p(bb), _p(out), len(aa))
    return [from_mont_limbs(r) for r in out]


def _is_lookup(c):
    return hasattr(c, "occurrences_id")


def c_cfgs(cfgs):
    """Registers the lookup configs of a `LineaAIR` config list with the library (they are folded first,
    trace/src/lib.rs:80-89) and returns the permutation configs as a C array.  Returns (array, keep-alive, log_q)."""
    lookups = [c for c in cfgs if _is_lookup(c)]
    cfgs = [c for c in cfgs if not _is_lookup(c)]
    blob = []
    for l in lookups:
        blob += [len(l.a_columns_ids), len(l.b_columns_ids), len(l.b_columns_ids[0]), l.a_filter_id, l.a_inverses_id, l.check_id]
        blob += list(l.a_columns_ids)
        for t, ids in enumerate(l.b_columns_ids):
            blob += [l.b_filter_id[t], l.b_inverses_id[t], l.occurrences_id[t]] + list(ids)
    barr = (C.c_uint32 * max(1, len(blob)))(*blob)
    if load().lsp_oracle_set_lookups(barr, len(lookups)) != 0:
        raise RuntimeError("lookup configs too large for the C oracle")
    keep, arr = [], (AirCfg * max(1, len(cfgs)))()
    keep.append(2 if lookups else 1)
    for i, c in enumerate(cfgs):
        a = (C.c_uint32 * len(c.a_columns_ids))(*c.a_columns_ids)
        b = (C.c_uint32 * len(c.b_columns_ids))(*c.b_columns_ids)
        keep += [a, b]
        arr[i] = AirCfg(len(c.a_columns_ids), a, b, c.b_inverse_id, c.check_id)
    keep.append(len(cfgs))
    return arr, keep


def proof_words(log_n, width, fri, log_q=1):
    f = FriCfg(fri.log_blowup, fri.log_final_poly_len, fri.num_queries, fri.proof_of_work_bits)
    return int(load().lsp_oracle_proof_words(log_n, width, log_q, C.byref(f)))


def prove_limbs(fri, trace_limbs: np.ndarray, n: int, w: int, cfgs, publics_limbs: np.ndarray, timings=None):
    """trace_limbs: uint64[n*w,4] row-major Montgomery.  Returns the flat proof (uint64 words),
    in the same layout as the CUDA library (linea-stark-prover_b200/host/prover.cu)."""
    f = FriCfg(fri.log_blowup, fri.log_final_poly_len, fri.num_queries, fri.proof_of_work_bits)
    arr, keep = c_cfgs(cfgs)
    words = proof_words(n.bit_length() - 1, w, fri, keep[0])
    out = np.zeros(words, dtype=np.uint64)
    tm = np.zeros(8, dtype=np.float64)
    rc = load().lsp_oracle_prove(C.byref(f), _p(trace_limbs), n, w, arr, keep[-1], _p(publics_limbs), _p(out), words,
                                 tm.ctypes.data_as(f64p))
    if rc != 0:
        raise RuntimeError(f"lsp_oracle_prove failed: {rc}")
    if timings is not None:
        timings[:] = tm
    del keep
    return out


def prove(fri, cfgs, trace, publics):
    n, w = len(trace), len(trace[0])
    return prove_limbs(fri, _arr([x for r in trace for x in r]), n, w, cfgs, _arr(publics))


def verify_limbs(fri, log_n: int, w: int, cfgs, publics_limbs: np.ndarray, proof_words_arr: np.ndarray) -> int:
    """0 = accepted; otherwise the failing check (see lsp_oracle_verify)."""
    f = FriCfg(fri.log_blowup, fri.log_final_poly_len, fri.num_queries, fri.proof_of_work_bits)
    arr, keep = c_cfgs(cfgs)
    rc = load().lsp_oracle_verify(C.byref(f), log_n, w, arr, keep[-1], _p(publics_limbs),
                                  _p(np.ascontiguousarray(proof_words_arr)), len(proof_words_arr))
    del keep
    return rc


def gen_trace(seed: int, c: int, log_n: int):
    """Synthetic permutation-argument witness (SURVEY.md 8(d) workload): returns
    (publics_limbs uint64[2,4], trace_limbs uint64[n*w,4], n, w)."""
    n, w = 1 << log_n, 2 * c + 2
    pub = np.zeros((2, 4), dtype=np.uint64)
    tr = np.zeros((n * w, 4), dtype=np.uint64)
    rc = load().lsp_oracle_gen_trace(seed, c, log_n, _p(pub), _p(tr))
    if rc != 0:
        raise RuntimeError(f"lsp_oracle_gen_trace failed: {rc}")
    return pub, tr, n, w


def permutation_trace(ab_limbs: np.ndarray, n: int, c: int, publics_limbs: np.ndarray) -> np.ndarray:
    """Witness of the permutation argument (trace/src/permutation.rs:24-93) from given a/b columns:
    ab_limbs uint64[n*2c,4] row-major Montgomery -> trace uint64[n*(2c+2),4]."""
    tr = np.zeros((n * (2 * c + 2), 4), dtype=np.uint64)
    rc = load().lsp_oracle_permutation_trace(_p(np.ascontiguousarray(ab_limbs)), n, c,
                                             _p(np.ascontiguousarray(publics_limbs)), _p(tr))
    if rc != 0:
        raise RuntimeError(f"lsp_oracle_permutation_trace failed: {rc} (b is not a permutation of a)")
    return tr


def set_field_consts(generator: int, two_adic_root: int):
    """Twin of `lsp_set_field_consts` (canonical ints in)."""
    if load().lsp_oracle_set_field_consts(_p(_arr([generator])), _p(_arr([two_adic_root]))) != 0:
        raise ValueError("bad field constants")


def set_transcript_flags(alpha_before_openings: bool = True, observe_opened_values: bool = False):
    load().lsp_oracle_set_transcript_flags(int(alpha_before_openings), int(observe_opened_values))


def threads() -> int:
    return int(load().lsp_oracle_threads())


def set_threads(n: int = 0) -> int:
    """n <= 0: every online processor, overriding OMP_NUM_THREADS (torchrun exports OMP_NUM_THREADS=1)."""
    return int(load().lsp_oracle_set_threads(int(n)))
