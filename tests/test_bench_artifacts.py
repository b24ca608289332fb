"""The committed bench lines of the round (profiles/r2_bench_*.json) must satisfy bench.py's output contract and be arithmetically
self-consistent: value = events / time, roofline.frac = achieved / peak, every HBM-table row = bytes / time / peak, e2e bytes
declared, clocks clean.  (CPU test: it reads the artefacts, it does not run the benchmark.)"""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROFILES = os.path.join(ROOT, "profiles")
BAD_REASONS = {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
CONTRACT = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
            "dtype", "data", "config", "e2e", "gpu_launches", "clocks"]


def _load(name):
    path = os.path.join(PROFILES, name)
    if not os.path.exists(path):
        pytest.skip(f"{name} not committed")
    return json.load(open(path))


@pytest.mark.parametrize("name,events_per_gpu", [("r2_bench_train.json", 16), ("r2_bench_infer.json", 32), ("r2_bench_gauge1pct.json", 16),
                                                  ("r2_bench_stress256.json", 8), ("r2_bench_train_2gpu.json", 16),
                                                  ("r2_bench_infer_8gpu.json", 32), ("r2_bench_train_8gpu_bucketed.json", 16)])
def test_bench_line_contract_and_arithmetic(name, events_per_gpu):
    d = _load(name)
    for k in CONTRACT:
        assert k in d, (name, k)
    assert d["unit"] == "events/s" and d["higher_is_better"] is True and d["scaling"] == "weak" and d["data"] == "synthetic"
    assert d["vs_baseline"] is None                       # BASELINE.md publishes no number for this metric
    assert d["warmup"] >= 3 and d["gpu_launches"] > 0
    assert d["config"]["events_per_step_per_gpu"] == events_per_gpu and "workload" in d["config"] and "model" not in d["config"]
    events = events_per_gpu * d["n_gpus"] * d["steps"]
    assert d["value"] == pytest.approx(events / (d["ms_per_step"] * d["steps"] * 1e-3), rel=1e-6)
    e = d["e2e"]
    assert e["unit"] == "events/s" and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert e["value"] != d["value"]                       # measured separately, not a copy of the device-resident number
    assert not (set(d["clocks"]["reasons"]) & BAD_REASONS), d["clocks"]
    assert d["clocks"]["sm_mhz"] >= 0.9 * d["clocks"]["sm_max_mhz"]
    r = d.get("roofline")
    if r:
        assert r["bound"] == "tensor" and r["unit"] == "TFLOP/s"
        assert r["frac"] == pytest.approx(r["achieved"] / r["peak"], rel=1e-6)
        assert r["achieved"] == pytest.approx(r["algorithmic_gflop_per_step"] / r["kernel_ms_per_step"], rel=1e-6)
        assert r["kernel_ms_per_step"] < d["ms_per_step"] * 1.02 or d["n_gpus"] > 1


def test_train_line_carries_hbm_table_baseline_and_subrecords():
    d = _load("r2_bench_train.json")
    r = d["roofline"]
    assert r["traffic"] and r["traffic"] > 0 and "ncu" in r["traffic_source"]
    hk = r["hbm_kernels"]
    assert isinstance(hk, list) and len(hk) >= 20
    for row in hk:
        gbs = row["algorithmic_mb_per_step"] * 1e-3 / (row["ms_per_step"] * 1e-3)
        assert row["achieved_gbs"] == pytest.approx(gbs, rel=1e-6), row
        assert row["frac_of_hbm"] == pytest.approx(row["achieved_gbs"] / r["hbm_peak_gbs"], rel=1e-6)
        assert 0 < row["frac_of_hbm"] < 1.0, row
    cb = d["cpu_baseline"]
    assert cb["unit"] == "events/s" and cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] > 0 and cb["sample"]
    for sub, cfg in (("infer", "configs[1]"), ("gauge1pct", "configs[3]"), ("stress256", "configs[4]")):
        assert d[sub]["baseline_config"] == cfg and d[sub]["value"] > 0 and d[sub]["e2e"]["value"] > 0
    sw = d["stress256"]["metrics_sweep"]
    assert sw["shape"] == [8, 20, 1, 256, 256] and 0 < sw["p2i_metrics_update"]["frac_of_hbm"] < 1
    assert d["losses_last_step"]["finite"] is True
    assert d["timed_region_s"] >= 2.0


def test_reference_arm_line():
    d = _load("r2_bench_reference_arm.json")
    assert d["impl"] == "reference" and d["unit"] == "events/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["value"] == pytest.approx(d["value"])
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
