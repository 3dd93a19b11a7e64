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
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--steps", "2", "--warmup", "3", "--log-n", "13", "--cfg3", "0"],
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
    assert rf["bound"].startswith("int32-multiply") and rf["peak"] > 0
    assert abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9 and 0 < rf["frac"] < 1.05
    assert rf["imad_wide_per_perm"] == 46 * 304 and len(rf["peak_forms_mac32_per_s"]) == 4
    assert abs(rf["peak"] * 1e12 - max(rf["peak_forms_mac32_per_s"].values())) < 1e-3 * rf["peak"] * 1e12
    assert abs(rf["hbm_frac"] - rf["hbm_achieved_gbs"] / rf["hbm_peak_gbs"]) < 1e-9
    assert abs(rf["hbm_achieved_gbs"] - rf["algorithmic_bytes_per_launch"] / (rf["ms_per_launch"] * 1e-3) / 1e9) < 1e-6 * rf["hbm_achieved_gbs"]
    assert 0 < rf["share_of_step"] < 1
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] > d["value"]      # the CPU port is slower than the GPU
    assert "scaled" not in cb["sample"].replace("nothing sampled or scaled", "")
    par = d["parity"]
    assert par["cpu_port_equal"] is True and par["witness_equal"] is True and par["e2e_proof_equal"] is True
    assert par["fnv1a64"] == par["fnv1a64_cpu_port"]
    e = d["e2e"]
    assert e["unit"] == "s" and e["value"] >= d["value"] * 0.9
    assert e["h2d_bytes_per_step"] == (1 << 13) * 8 * 32 and e["d2h_bytes_per_step"] > 0
    assert d["gpu_launches"] > 50
    assert d["clocks"]["sm_max_mhz"] >= d["clocks"]["sm_mhz"] > 0 and isinstance(d["clocks"]["reasons"], list)
    assert d["verify"]["accepted"] is True and d["verify"]["device_ms"] > 0
