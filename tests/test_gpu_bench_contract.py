"""bench.py's own arm at a small size on the GPU: one JSON line with every object the contract names
(`roofline`, `cpu_baseline`, `e2e`, `gpu_launches`, `clocks`), and numbers that are internally consistent."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def test_our_arm_prints_one_json_line_with_every_contract_object():
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--steps", "2", "--warmup", "3", "--log-n", "13", "--cpu-budget", "1"],
                       capture_output=True, text=True, timeout=900, cwd=str(ROOT))
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout[-2000:]
    d = json.loads(lines[0])
    assert d["metric"] == "prove_seconds" and d["unit"] == "s" and d["higher_is_better"] is False
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 3 and d["data"] == "synthetic" and d["vs_baseline"] is None
    assert d["value"] > 0 and abs(d["ms_per_step"] - d["value"] * 1e3) < 1e-6
    assert d["config"]["rows"] == 1 << 13 and "l2" in d["config"]
    rf = d["roofline"]
    assert rf["bound"] == "hbm" and rf["unit"] == "GB/s" and rf["peak"] > 0
    assert abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9
    assert abs(rf["achieved"] - rf["algorithmic_bytes_per_launch"] / (rf["ms_per_launch"] * 1e-3) / 1e9) < 1e-6 * rf["achieved"]
    assert 0 < rf["share_of_step"] < 1
    ir = d["int_roofline"]
    assert 0 < ir["frac"] < 1.05 and ir["imad_wide_per_perm"] == 46 * 304
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] > d["value"]      # the CPU port is slower than the GPU
    e = d["e2e"]
    assert e["unit"] == "s" and e["value"] >= d["value"] * 0.9
    assert e["h2d_bytes_per_step"] == (1 << 13) * 8 * 32 and e["d2h_bytes_per_step"] > 0
    assert d["gpu_launches"] > 50
    assert d["clocks"]["sm_max_mhz"] >= d["clocks"]["sm_mhz"] > 0 and isinstance(d["clocks"]["reasons"], list)
    assert d["verify"]["accepted"] is True and d["verify"]["device_ms"] > 0
