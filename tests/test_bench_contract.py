"""CPU-only check of the bench.py output contract on the committed line of the last GPU run
(profiles/bench_r1_final.json): every key the driver reads is present and self-consistent, and the reference arm
line has its own required keys.  (The line itself is produced on a B200; this test only guards the schema.)"""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load(name):
    with open(os.path.join(ROOT, "profiles", name)) as f:
        return json.loads(f.read().strip().splitlines()[-1])


def test_main_line_has_the_contract_keys():
    d = load("bench_r1_final.json")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "cpu_baseline", "clocks"):
        assert k in d, k
    assert d["n_gpus"] == 1 and d["warmup"] >= 3 and d["higher_is_better"] is True and d["scaling"] == "weak"
    assert d["vs_baseline"] is None                       # BASELINE.md publishes no number for this metric on this hardware
    assert "workload" in d["config"] and "l2" in d["config"]
    assert abs(d["value"] - d["config"]["pairs_total"] / (d["ms_per_step"] * 1e-3)) / d["value"] < 1e-6
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] == 192 * d["config"]["pairs_total"] and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] < d["value"]
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["unit"] == "GB/s"
    assert r["traffic"] is None or r["traffic"] > 0
    c = d["cpu_baseline"]
    assert c["kind"] in ("port", "reference") and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
    assert d["gpu_launches"] > 0
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert d["clocks"]["sm_mhz"] > 0.9 * d["clocks"]["sm_max_mhz"]
    # device-event timing agrees with the bracketed wall clock
    assert abs(d["config"]["device_ms_per_step"] - d["ms_per_step"]) / d["ms_per_step"] < 0.05
    v = d["verify_batch"]
    assert v["unit"] == "sigs/s" and v["e2e"]["h2d_bytes_per_step"] > 155 * v["config"]["signatures_per_gpu"]


def test_reference_arm_line():
    d = load("bench_r1_reference_arm.json")
    assert d["impl"] == "reference" and d["metric"] == load("bench_r1_final.json")["metric"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["cores"] >= 1
