"""The bench line's contract, checked on the last committed measurement (profiles/): every key the
driver reads is there, with the types and relations the contract states."""
import glob
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _latest(pattern):
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", pattern)), key=os.path.getmtime)
    assert files, pattern
    return json.loads(open(files[-1]).read())


def test_own_arm_line():
    d = _latest("r01zx_bench.json")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert k in d, k
    assert d["unit"] == "Mpix/s" and d["dtype"] == "f32" and d["data"] == "synthetic" and d["vs_baseline"] is None
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and "workload" in d["config"]
    assert "model" not in d["config"] and "l2" in d["config"]
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] in ("GB/s", "TFLOP/s")
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["traffic"] > 0
    c = d["cpu_baseline"]
    assert c["kind"] in ("reference", "port") and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] < d["value"]
    assert d["gpu_launches"] > 0
    assert d["clocks"]["sm_mhz"] > 0 and not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown",
                                                                              "sw_thermal_slowdown"}
    # value is pixels of the whole job over the step time
    cfg = d["config"]
    pix = d["n_gpus"] * cfg["hypotheses_per_gpu"] * cfg["width"] * cfg["height"]
    assert abs(d["value"] - pix / (d["ms_per_step"] * 1e-3) / 1e6) < 1e-6 * d["value"]


def test_reference_arm_line():
    d = _latest("r01zv_bench_ref.json")
    assert d["impl"] == "reference" and d["unit"] == "Mpix/s" and d["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["kind"] in ("reference", "port")
