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
