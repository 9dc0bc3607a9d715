"""The bench.py contract, checked without a GPU: the reference arm runs here (CPU) and must print exactly ONE JSON line
with the agreed keys; the own arm's last committed line (profiles/) must carry roofline / cpu_baseline / e2e / clocks /
gpu_launches with consistent arithmetic."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT,
                         env={**__import__("os").environ, "CRDPN_REF_SAMPLE_B": "2"})
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, out.stdout[:500]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "crd_negatives_scored_per_sec" and d["unit"] == "scores/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0
    assert d["value"] > 0 and d["config"]["workload"].startswith("crd_B46_D128_K65536_N1000000")
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "scores/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_committed_own_arm_line_has_the_contract_keys():
    d = json.loads((ROOT / "profiles" / "r1_bench_1gpu_final.json").read_text().strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert key in d, key
    assert d["vs_baseline"] is None and d["dtype"] == "f32" and d["data"] == "synthetic" and "workload" in d["config"]
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert abs(r["achieved"] - r["algorithmic_bytes"] / (r["kernel_ms"] * 1e-3) / 1e9) < 1e-6 * r["achieved"]
    assert r["traffic"] is None or r["traffic"] > 0
    scores = 2 * d["config"]["B"] * (d["config"]["K"] + 1)
    assert abs(d["value"] - scores / (d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"]
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] < d["value"]
    assert d["gpu_launches"] >= d["steps"] and set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    c = d["cpu_baseline"]
    assert c["kind"] in ("port", "reference") and c["cores"] >= 1 and c["sample"] and c["unit"] == "scores/s"
