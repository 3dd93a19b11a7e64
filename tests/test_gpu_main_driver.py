"""`lsp_prove`, the C++ restatement of the reference's `main` (bin/src/main.rs:19-97) over the C ABI: run as a separate
process on CBOR input files, its proof must be byte-identical to the one the Python mirror produces (and that the
oracle's verifier accepts) for the same seed."""
import subprocess
from pathlib import Path

import numpy as np
import pytest

from oracle import field as F
from oracle import stark as OS
from oracle import trace as OT
from oracle.poseidon2 import Poseidon2Params

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent
EXE = ROOT / "linea-stark-prover_b200" / "lsp_prove"


def _draws(seed):
    """The driver's seeded draws: limbs below r taken as MONTGOMERY representatives (any such value is one)."""
    rng = F.SplitMix64(seed)
    nxt = lambda: F.from_mont(rng.next_fr())
    alpha, delta = nxt(), nxt()
    consts = [nxt() for _ in range(8 * 3 + 22)]
    return alpha, delta, consts


def test_driver_proves_cbor_files_like_main(pkg, tmp_path):
    assert EXE.exists(), "lsp_prove is built by `make -C linea-stark-prover_b200` / __graft_entry__.build()"
    n, seed = 64, 4242
    alpha, delta, consts = _draws(seed)
    lk = OT.synthetic_lookup_input(1, 2, 2, n, disabled_every=9)
    pa, pb = OT.synthetic_permutation_input(2, 3, n)
    (tmp_path / "lookup_0.bin").write_bytes(OT.encode_raw_lookup_trace(*lk, "lookup_0"))
    (tmp_path / "perm_0.bin").write_bytes(OT.encode_raw_permutation_trace(pa, pb, "perm_0"))
    out, ser = tmp_path / "proof.bin", tmp_path / "proof.lspp"
    r = subprocess.run([str(EXE), "--lookup", str(tmp_path / "lookup_0.bin"), "--permutation", str(tmp_path / "perm_0.bin"), "--seed",
                        str(seed), "--queries", "9", "--out", str(out), "--out-serialized", str(ser)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "Proving..." in r.stdout and "commit to trace data" in r.stdout
    assert "Verifying..." in r.stdout and "proof accepted" in r.stdout          # main.rs:88-96
    words = np.frombuffer(out.read_bytes(), dtype=np.uint64)
    # the same proof through the Python mirror, and the oracle accepts it
    p = Poseidon2Params(sbox_d=5, rounds_f=8, rounds_p=22, ext_initial=[consts[3 * i:3 * i + 3] for i in range(4)],
                        ext_terminal=[consts[12 + 3 * i:15 + 3 * i] for i in range(4)], internal=consts[24:], internal_diag_m1=(1, 1, 2))
    ctx = pkg.Context(0)
    ctx.set_poseidon2(5, 8, 22, p.flat_constants(), p.internal_diag_m1)
    cfgs, trace = OT.build_trace([(pa, pb)], alpha, delta, [lk])
    from tests.test_gpu_lookup import _gpu_cfgs
    fri = dict(log_blowup=3, log_final_poly_len=0, num_queries=9, proof_of_work_bits=0)
    mine = pkg.prove(ctx, pkg.FriConfig(**fri), _gpu_cfgs(pkg, cfgs), trace, [alpha, delta])
    assert np.array_equal(words, mine.words)
    gd, _ = mine.to_dict()
    OS.verify(p, OS.FriConfig(**fri), cfgs, gd, [alpha, delta])
    # the serialised form the driver wrote is the mirror's, and what it carries verifies on the device and in the oracle
    assert ser.read_bytes() == mine.serialize()
    back = pkg.Proof.deserialize(ser.read_bytes())
    pkg.verify(ctx, pkg.FriConfig(**fri), _gpu_cfgs(pkg, cfgs), back, [alpha, delta])
    assert back.to_dict()[0] == gd
    ctx.close()


def test_driver_pads_inputs_of_different_heights_like_push_traces(pkg, tmp_path):
    """`RawTrace::push_traces` (trace/src/lib.rs:62-92): the height is the tallest column of any input file and every
    sub-trace is zero-padded to it before its witness is built -- a 16-row lookup beside a 64-row permutation proves."""
    seed = 99
    alpha, delta, consts = _draws(seed)
    lk = OT.synthetic_lookup_input(7, 2, 1, 16, disabled_every=5)
    pa, pb = OT.synthetic_permutation_input(8, 2, 64)
    pa2, pb2 = OT.synthetic_permutation_input(9, 1, 32)
    (tmp_path / "lookup_0.bin").write_bytes(OT.encode_raw_lookup_trace(*lk, "lookup_0"))
    (tmp_path / "perm_0.bin").write_bytes(OT.encode_raw_permutation_trace(pa, pb, "perm_0"))
    (tmp_path / "perm_1.bin").write_bytes(OT.encode_raw_permutation_trace(pa2, pb2, "perm_1"))
    out = tmp_path / "proof.bin"
    r = subprocess.run([str(EXE), "--lookup", str(tmp_path / "lookup_0.bin"), "--permutation", str(tmp_path / "perm_0.bin"),
                        "--permutation", str(tmp_path / "perm_1.bin"), "--seed", str(seed), "--queries", "7", "--out", str(out)],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "proof accepted" in r.stdout and "64 rows" in r.stdout
    words = np.frombuffer(out.read_bytes(), dtype=np.uint64)
    p = Poseidon2Params(sbox_d=5, rounds_f=8, rounds_p=22, ext_initial=[consts[3 * i:3 * i + 3] for i in range(4)],
                        ext_terminal=[consts[12 + 3 * i:15 + 3 * i] for i in range(4)], internal=consts[24:], internal_diag_m1=(1, 1, 2))
    ctx = pkg.Context(0)
    ctx.set_poseidon2(5, 8, 22, p.flat_constants(), p.internal_diag_m1)
    cfgs, trace = OT.build_trace([(pa, pb), (pa2, pb2)], alpha, delta, [lk])
    assert len(trace) == 64
    from tests.test_gpu_lookup import _gpu_cfgs
    fri = dict(log_blowup=3, log_final_poly_len=0, num_queries=7, proof_of_work_bits=0)
    mine = pkg.prove(ctx, pkg.FriConfig(**fri), _gpu_cfgs(pkg, cfgs), trace, [alpha, delta])
    assert np.array_equal(words, mine.words)
    gd, _ = mine.to_dict()
    OS.verify(p, OS.FriConfig(**fri), cfgs, gd, [alpha, delta])
    ctx.close()


def test_driver_fails_loudly_on_bad_input(tmp_path):
    (tmp_path / "junk.bin").write_bytes(b"\x00\x01\x02")
    r = subprocess.run([str(EXE), "--permutation", str(tmp_path / "junk.bin")], capture_output=True, text=True, timeout=120)
    assert r.returncode != 0 and "lsp_cbor_permutation_read" in r.stderr


def test_driver_shards_one_proof_over_two_gpus(tmp_path):
    """`lsp_prove --gpus 2`: one host thread and one context per GPU, NCCL between them; the proof must be the
    single-GPU proof byte for byte (and the binary itself verifies it on the device before writing it)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    n, seed = 256, 777
    lk = OT.synthetic_lookup_input(3, 2, 1, n, disabled_every=5)
    pa, pb = OT.synthetic_permutation_input(4, 2, n)
    (tmp_path / "lookup_0.bin").write_bytes(OT.encode_raw_lookup_trace(*lk, "lookup_0"))
    (tmp_path / "perm_0.bin").write_bytes(OT.encode_raw_permutation_trace(pa, pb, "perm_0"))
    outs = []
    for gpus in (1, 2):
        out = tmp_path / f"proof_{gpus}.bin"
        r = subprocess.run([str(EXE), "--lookup", str(tmp_path / "lookup_0.bin"), "--permutation", str(tmp_path / "perm_0.bin"),
                            "--seed", str(seed), "--queries", "12", "--gpus", str(gpus), "--out", str(out)],
                           capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout + r.stderr
        assert "proof accepted" in r.stdout
        outs.append(out.read_bytes())
    assert outs[0] == outs[1] and len(outs[0]) > 0


def test_driver_rejects_bad_gpu_counts(tmp_path):
    (tmp_path / "p.bin").write_bytes(OT.encode_raw_permutation_trace(*OT.synthetic_permutation_input(1, 1, 8), "p"))
    for gpus in ("3", "16", "0"):
        r = subprocess.run([str(EXE), "--permutation", str(tmp_path / "p.bin"), "--gpus", gpus], capture_output=True, text=True, timeout=120)
        assert r.returncode == 2 and "--gpus" in r.stderr
