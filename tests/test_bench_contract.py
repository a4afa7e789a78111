"""CPU: the committed bench lines (profiles/r2_*.json, produced by bench.py on the B200) carry every key the measurement
contract names, with self-consistent values -- a guard against bench.py edits that silently drop a field."""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(name):
    path = os.path.join(ROOT, "profiles", name)
    if not os.path.exists(path):
        pytest.skip(f"{name} not committed")
    return json.loads(open(path).read().strip().splitlines()[-1])


BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config", "e2e", "gpu_launches", "clocks"}


def test_single_gpu_line_has_the_contract_keys_and_is_self_consistent():
    d = _line("r2_bench_ours.json")
    assert BASE_KEYS <= set(d), BASE_KEYS - set(d)
    assert d["n_gpus"] == 1 and d["higher_is_better"] is True and d["vs_baseline"] is None and "workload" in d["config"]
    # tok/s = batch / step time
    assert abs(d["value"] - d["config"]["global_batch"] / (d["ms_per_step"] * 1e-3)) / d["value"] < 1e-3
    e = d["e2e"]
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(e) and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert e["value"] < d["value"]                       # copies inside the timed region cannot make it faster
    r = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and r["bound"] in ("hbm", "tensor")
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-6 and 0 < r["frac"] < 1
    k = d["kernels"]["w4a8_gemm(decode,4 launches/layer)"]
    assert abs(r["achieved"] - k["algorithmic_bytes"] / (k["ms_per_layer"] * 1e-3) / 1e9) / r["achieved"] < 1e-3
    assert r["traffic"] is None or 0.5 < r["traffic"] / k["algorithmic_bytes"] < 1.5     # no wasted re-reads
    c = d["cpu_baseline"]
    assert {"value", "unit", "cores", "kind", "sample"} <= set(c) and c["kind"] in ("port", "reference") and c["cores"] >= 1
    assert d["clocks"]["sm_mhz"] > 0.9 * d["clocks"]["sm_max_mhz"]
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert d["gpu_launches"] > 0


def test_reference_arm_line():
    d = _line("r2_bench_reference.json")
    assert d.get("impl") == "reference" and BASE_KEYS - {"gpu_launches"} <= set(d)
    ours = _line("r2_bench_ours.json")
    assert d["metric"] == ours["metric"] and d["unit"] == ours["unit"] and d["config"]["workload"] == ours["config"]["workload"]


def test_tensor_parallel_line_reports_strong_scaling_with_dp_and_c4_secondaries():
    d = _line(os.path.join("r2_tp", "bench_tp8.json"))
    assert d["n_gpus"] == 8 and d["scaling"] == "strong" and d["config"]["parallelism"] == "tp8"
    assert d["dp"]["parallelism"] == "dp8" and d["dp"]["scaling"] == "weak" and d["dp"]["value"] > d["value"]
    assert "70B" in d["c4"]["model"] and d["c4"]["global_batch"] == 16 and d["c4"]["value"] > 0
