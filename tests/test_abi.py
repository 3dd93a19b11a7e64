"""The C-ABI library loads and exports every symbol include/lsp_b200.h declares
(no compute calls: this runs without a GPU)."""
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "lsp_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lsp_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(pkg):
    lib = pkg.ffi.load()
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in lsp_b200.h but not exported"
        assert s in pkg.ffi.SIGNATURES, f"{s} has no ctypes signature"
    assert lib.lsp_abi_version() == 1


def test_no_cpu_fallback_without_device(pkg):
    """Without a CUDA device the backend must fail loudly, not fall back."""
    import torch
    if torch.cuda.is_available():
        return
    import pytest
    with pytest.raises(pkg.BackendError):
        pkg.Context(0)


def test_product_never_imports_oracle():
    """oracle/ is the checker; nothing under the product tree may import, include or dlopen it."""
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|#include\s+[\"<][^\">]*oracle|dlopen\([^)]*oracle", re.M)
    for path in (ROOT / "linea-stark-prover_b200").rglob("*"):
        if path.suffix in {".py", ".cu", ".cuh", ".cpp", ".hpp", ".h"}:
            assert not pat.search(path.read_text()), path


def test_quotient_degree_is_derived_like_the_symbolic_pass(pkg):
    """`get_log_quotient_degree`: the library derives it from `LineaAIR::eval` on symbolic degrees (host only,
    no GPU), as the oracle does -- not from a table keyed on "has a lookup"."""
    import ctypes as C

    from oracle import air as OA
    lib = pkg.ffi.load()
    cases = [[OA.AirPermutationConfig.standard(1)], [OA.AirPermutationConfig.standard(6)],
             [OA.AirLookupConfig.standard(1, 1, 1)], [OA.AirLookupConfig.standard(2, 3, 2)],
             [OA.AirLookupConfig.standard(2, 2, 2), OA.AirPermutationConfig.standard(3)]]
    for cfgs in cases:
        lk = [c for c in cfgs if isinstance(c, OA.AirLookupConfig)]
        pm = [c for c in cfgs if not isinstance(c, OA.AirLookupConfig)]
        larr = (pkg.ffi.LookupAirCfg * max(1, len(lk)))()
        parr = (pkg.ffi.PermAirCfg * max(1, len(pm)))()
        for i, c in enumerate(lk):
            larr[i].n_a_cols, larr[i].n_tables, larr[i].n_b_cols = len(c.a_columns_ids), len(c.b_columns_ids), len(c.b_columns_ids[0])
        for i, c in enumerate(pm):
            parr[i].n_cols = len(c.a_columns_ids)
        got = lib.lsp_air_log_quotient_degree_cfg(larr, len(lk), parr, len(pm))
        assert got == OA.log_quotient_degree(cfgs) == lib.lsp_air_log_quotient_degree(len(lk), len(pm)), cfgs
    assert lib.lsp_air_log_quotient_degree_cfg(None, 1, None, 0) < 0          # a count without its configs is an error


def test_rust_bindings_declare_existing_symbols_with_the_headers_arity():
    """`ffi/lsp-b200-sys/src/lib.rs` cannot be compiled here (no rustc): at least every `pub fn lsp_*` it declares must exist
    in `include/lsp_b200.h` with the same number of parameters, and the repr(C) structs must list the header's fields in order."""
    import re
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    header = re.sub(r"/\*.*?\*/", "", (root / "include" / "lsp_b200.h").read_text(), flags=re.S)
    rust = re.sub(r"//.*", "", (root / "ffi" / "lsp-b200-sys" / "src" / "lib.rs").read_text())

    def arity(params):
        params = params.strip()
        return 0 if params in ("", "void") else len([p for p in re.split(r",(?![^\[]*\])", params) if p.strip()])

    c_fns = {m.group(1): arity(m.group(2)) for m in re.finditer(r"\b(lsp_\w+)\s*\(([^;{]*?)\)\s*;", header, flags=re.S)}
    r_fns = {m.group(1): arity(m.group(2)) for m in re.finditer(r"pub fn (lsp_\w+)\s*\((.*?)\)\s*(?:->[^;]+)?;", rust, flags=re.S)}
    assert len(r_fns) >= 40
    for name, n in r_fns.items():
        assert name in c_fns, f"{name} is not declared in include/lsp_b200.h"
        assert c_fns[name] == n, f"{name}: {n} parameters in the Rust binding, {c_fns[name]} in the header"
    # struct field order
    for c_name in ("lsp_fri_config", "lsp_perm_air_cfg", "lsp_lookup_air_cfg"):
        c_body = re.search(r"typedef struct \{([^{}]*)\}\s*" + c_name + r"\s*;", header).group(1)
        c_fields = [re.findall(r"(\w+)\s*$", f.strip())[0] for f in c_body.split(";") if f.strip()]
        r_body = re.search(r"pub struct " + c_name + r"\s*\{(.*?)\}", rust, flags=re.S).group(1)
        r_fields = re.findall(r"pub (\w+)\s*:", r_body)
        assert c_fields == r_fields, (c_name, c_fields, r_fields)
