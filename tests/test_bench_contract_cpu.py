"""CPU: the committed bench lines (profiles/r2_bench_final.json, r2_bench_reference_final.json, r2_bench_n8.json — written by
bench.py on B200s) carry every key the measurement contract names, with consistent values."""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(name):
    return json.loads(open(os.path.join(ROOT, "profiles", name)).read().strip().splitlines()[-1])


@pytest.mark.parametrize("name", ["r2_bench_final.json", "r2b_bench_final.json"])
def test_ours_line_has_the_contract_keys(name):
    d = _line(name)
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline", "mapping",
              "batch", "configs"):
        assert k in d, k
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert d["metric"] == base["metric"].replace("\u00d7", "x").replace("×", "x") or d["metric"] == base["metric"]
    assert d["n_gpus"] == 1 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["warmup"] == 5 and d["steps"] == 20                       # --warmup / --steps are honoured
    assert d["timed_region_s"] >= 1.0                                  # >= 1 s inside the timed region
    assert d["dtype"] == "f32" and d["data"] == "synthetic" and "workload" in d["config"] and "model" not in d["config"]
    assert d["config"]["views_in_flight_per_gpu"] == 1                 # `value` is M1: one view at a time
    assert d["gpu_launches"] > 0
    assert set(("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step")) <= set(d["e2e"])
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0 and d["e2e"]["value"] != d["value"]
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor", "issue") and r["unit"] in ("GB/s", "TFLOP/s", "Gwarp-inst/s")
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-3 and r["traffic"] is not None
    h = r["hbm"]                                                       # the contract's byte formula, restated
    assert h["unit"] == "GB/s" and abs(h["frac"] - h["achieved"] / h["peak"]) < 1e-3
    assert abs(h["achieved"] - h["algorithmic_MB"] / r["ms_per_launch"]) / h["achieved"] < 2e-3
    c = d["cpu_baseline"]
    assert c["kind"] in ("reference", "port") and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
    assert set(("sm_mhz", "sm_max_mhz", "reasons")) <= set(d["clocks"])
    assert not any("slowdown" in x for x in d["clocks"]["reasons"])
    # value = views / time: 64 views per step and GPU
    assert abs(d["value"] - 64 * d["steps"] / (d["ms_per_step"] * d["steps"] * 1e-3)) / d["value"] < 1e-3
    m = d["mapping"]
    assert m["unit"] == "keyframes/s" and m["steps"] >= 20 and abs(m["value"] - 64 / (m["ms_per_step"] * 1e-3)) / m["value"] < 1e-3
    assert set(("min", "median", "max")) <= set(m["ms_per_step_rank0"]) and m["with_frequency"]["value"] > 0
    assert set(("C1", "C3", "C5", "distCUDA2")) <= set(d["configs"])
    assert d["batch"]["views_in_flight_per_gpu"] > 1 and d["batch"]["value"] > d["value"]


@pytest.mark.parametrize("ours,ref", [("r2_bench_final.json", "r2_bench_reference_final.json"),
                                      ("r2b_bench_final.json", "r2b_bench_reference_final.json")])
def test_reference_line_is_marked_and_comparable(ours, ref):
    o, r = _line(ours), _line(ref)
    assert r["impl"] == "reference" and r["gpu_launches"] == 0
    for k in ("metric", "unit", "higher_is_better", "steps", "warmup"):
        assert r[k] == o[k]
    assert r["config"]["workload"] == o["config"]["workload"]          # same config: M1 in both arms
    assert r["config"]["views_per_step_per_gpu"] == o["config"]["views_per_step_per_gpu"]
    assert r["config"]["views_in_flight_per_gpu"] == o["config"]["views_in_flight_per_gpu"] == 1
    assert r["e2e"]["h2d_bytes_per_step"] == o["e2e"]["h2d_bytes_per_step"]
    assert o["value"] > 3 * r["value"]                      # north_star: >= 3x the reference rasterizer on one B200
    assert o["e2e"]["value"] > 3 * r["e2e"]["value"]
    assert r["mapping"]["baseline_B"]["value"] > 0 and o["mapping"]["value"] > 3 * r["mapping"]["baseline_B"]["value"]
    for k in ("C1", "C3", "C5"):
        assert o["configs"][k]["iterations_per_s"] > 3 * r["configs"][k]["iterations_per_s"], k


def test_eight_gpu_line_meets_the_scaling_target():
    o, e = _line("r2_bench_final.json"), _line("r2_bench_n8.json")
    assert e["n_gpus"] == 8 and e["scaling"] == "weak"
    assert e["value"] / (8 * o["value"]) >= 0.85                        # north_star: >= 0.85 at 8 GPUs
    assert e["mapping"]["value"] / (8 * o["mapping"]["value"]) >= 0.85


def test_decode_variant_2_is_the_default_and_not_slower():
    """Round 2b: the C3 sub-line is measured with the default decode kernels (variant 2: both MLP layers and the weight
    gradients on tcgen05) and, beside it, with round 1's (variant 1); the mapping step moved with them."""
    d, before = _line("r2b_bench_final.json"), _line("r2_bench_final.json")
    c3 = d["configs"]["C3"]
    assert c3["decode_variant"] == 2 and c3["mean_ms"] < c3["decode_variant_1"]["mean_ms"]
    assert d["mapping"]["value"] > before["mapping"]["value"]
