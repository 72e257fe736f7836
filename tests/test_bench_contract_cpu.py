"""CPU: the committed bench lines (profiles/r1_bench_final.json, r1_bench_reference_final.json — written by bench.py on a
B200) carry every key the measurement contract names, with consistent values."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(name):
    return json.loads(open(os.path.join(ROOT, "profiles", name)).read().strip().splitlines()[-1])


def test_ours_line_has_the_contract_keys():
    d = _line("r1_bench_final.json")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline", "mapping"):
        assert k in d, k
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert d["metric"] == base["metric"].replace("×", "x") or d["metric"] == base["metric"]
    assert d["n_gpus"] == 1 and d["warmup"] >= 3 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["dtype"] == "f32" and d["data"] == "synthetic" and "workload" in d["config"] and "model" not in d["config"]
    assert d["gpu_launches"] > 0
    assert set(("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step")) <= set(d["e2e"])
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0 and d["e2e"]["value"] != d["value"]
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] in ("GB/s", "TFLOP/s")
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-3 and r["traffic"] is not None
    c = d["cpu_baseline"]
    assert c["kind"] in ("reference", "port") and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
    assert set(("sm_mhz", "sm_max_mhz", "reasons")) <= set(d["clocks"])
    assert not any("slowdown" in x for x in d["clocks"]["reasons"])
    # value = views / time: 8 views per step
    assert abs(d["value"] - 8 * d["steps"] / (d["ms_per_step"] * d["steps"] * 1e-3)) / d["value"] < 1e-3
    m = d["mapping"]
    assert m["unit"] == "keyframes/s" and abs(m["value"] - 64 / (m["ms_per_step"] * 1e-3)) / m["value"] < 1e-3


def test_reference_line_is_marked_and_comparable():
    o, r = _line("r1_bench_final.json"), _line("r1_bench_reference_final.json")
    assert r["impl"] == "reference" and r["gpu_launches"] == 0
    for k in ("metric", "unit", "higher_is_better"):
        assert r[k] == o[k]
    assert r["config"]["workload"] == o["config"]["workload"]
    assert r["e2e"]["h2d_bytes_per_step"] == o["e2e"]["h2d_bytes_per_step"]
    assert o["value"] > 3 * r["value"]                      # north_star: >= 3x the reference rasterizer on one B200
    assert o["e2e"]["value"] > 3 * r["e2e"]["value"]
