"""bench.py's output contract, checked on the arm that runs without a GPU (`--impl reference`: the C port of the
reference's CPU prover on the host cores): exactly ONE JSON line on stdout, the keys the driver reads, the same
`config` object our own arm prints, and -- under a multi-rank launch -- only rank 0 speaking."""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
ARGS = ["--impl", "reference", "--steps", "1", "--warmup", "0", "--log-n", "10"]


def _run(extra_env=None, extra_args=()):
    env = dict(os.environ)
    env.update(extra_env or {})
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), *ARGS, *extra_args], capture_output=True, text=True, timeout=600,
                       env=env, cwd=str(ROOT))
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    out = _run()
    lines = [l for l in out.splitlines() if l.strip()]
    assert len(lines) == 1, out
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "prove_seconds" and d["unit"] == "s"
    assert d["higher_is_better"] is False and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0
    assert d["value"] > 0 and abs(d["ms_per_step"] - d["value"] * 1e3) < 1e-6
    assert set(d["config"]) >= {"workload", "rows", "width", "log_blowup"} and "model" not in d["config"]
    assert d["config"]["rows"] == 1 << 10 and d["config"]["width"] == 8
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["unit"] == "s" and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # the arm proves the workload it names, in full: nothing sampled, nothing scaled
    assert "scaled" not in cb["sample"] and "2^10-row" in cb["sample"]
    assert len(d["parity"]["fnv1a64"]) == 16 and d["parity"]["verified_by_cpu_port"] is True


def test_reference_arm_proves_the_same_seeded_trace_every_time():
    a, b = json.loads(_run().strip()), json.loads(_run().strip())
    assert a["parity"]["fnv1a64"] == b["parity"]["fnv1a64"]


def test_reference_arm_config_equals_our_arms_config():
    import bench
    ap_args = type("A", (), dict(cols=3, log_n=10, log_blowup=3, sbox_d=5))()
    d = json.loads(_run().strip().splitlines()[-1])
    assert d["config"] == bench.workload_config(ap_args, 1)


def test_reference_arm_under_a_multi_rank_launch_only_rank_0_speaks():
    env = {"WORLD_SIZE": "2", "LOCAL_RANK": "1", "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": "29533"}
    assert _run(dict(env, RANK="1"), ("--gpus", "2")).strip() == ""
    d = json.loads(_run(dict(env, RANK="0", LOCAL_RANK="0"), ("--gpus", "2")).strip())
    assert d["impl"] == "reference" and d["n_gpus"] == 2
