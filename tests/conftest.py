import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def pytest_sessionstart(session):
    """A fresh checkout has no built artefacts (they are git-ignored): build them once, exactly as
    `__graft_entry__.build()` does, instead of failing every test that loads the library or the C oracle.
    Nothing is rebuilt when both shared objects are already there (the GPU box receives them prebuilt)."""
    lib = ROOT / "linea-stark-prover_b200" / "liblsp_b200.so"
    orc = ROOT / "oracle" / "c" / "liblsp_oracle.so"
    exe = ROOT / "linea-stark-prover_b200" / "lsp_prove"
    if lib.exists() and orc.exists() and exe.exists():
        return
    import shutil
    if shutil.which("nvcc") is None and not Path("/usr/local/cuda/bin/nvcc").exists():
        return      # nothing to build with: the tests that need the library will say so
    import __graft_entry__ as g
    g.build()


@pytest.fixture(scope="session")
def pkg():
    import __graft_entry__ as g
    return g.load_package()


@pytest.fixture(scope="session")
def p2params():
    from oracle.poseidon2 import Poseidon2Params
    return Poseidon2Params.from_seed(0xB200, sbox_d=5)


@pytest.fixture(scope="session")
def gctx(pkg, p2params):
    """A GPU context with the session's Poseidon2 constants installed.  Fails loudly
    (no skip) when the library or the device is missing: the product has no CPU path."""
    ctx = pkg.Context(0)
    ctx.set_poseidon2(p2params.sbox_d, p2params.rounds_f, p2params.rounds_p, p2params.flat_constants(),
                      p2params.internal_diag_m1)
    yield ctx
    ctx.close()
