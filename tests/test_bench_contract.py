"""The bench lines committed under profiles/ (produced by bench.py on a B200, tools/r2_final.sh) carry every key of the
measurement contract -- guards bench.py's output format without a GPU."""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(name):
    path = os.path.join(ROOT, "profiles", name)
    if not os.path.exists(path):
        pytest.skip(f"{name} not committed")
    return json.loads(open(path).read().strip().splitlines()[-1])


BASE_KEYS = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches"]


@pytest.mark.parametrize("name", ["r02_bench_c2.json", "r02_bench_c3.json", "r02_bench_c4.json"])
def test_single_gpu_lines(name):
    d = _line(name)
    for k in BASE_KEYS + ["roofline"]:
        assert k in d, k
    assert d["n_gpus"] == 1 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["gpu_launches"] > 0
    for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert k in d["e2e"], k
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["value"] != d["value"]
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    c = d["clocks"]
    assert c["sm_mhz"] and c["sm_max_mhz"] and isinstance(c["reasons"], list)
    assert not {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(c["reasons"])
    assert abs(d["value"] - d["config"]["sizes"]["G"] / (d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"]


def test_default_line_has_cpu_baseline():
    d = _line("r02_bench_c2.json")
    b = d["cpu_baseline"]
    for k in ("value", "unit", "cores", "kind", "sample"):
        assert k in b, k
    assert b["kind"] in ("port", "reference") and b["cores"] >= 1 and b["unit"] == d["unit"]


def test_reference_arm_line():
    d = _line("r02_bench_reference_arm.json")
    assert d["impl"] == "reference"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["e2e"]["value"] == d["value"] == d["cpu_baseline"]["value"]
    ours = _line("r02_bench_c2.json")
    assert d["metric"] == ours["metric"] and d["unit"] == ours["unit"]


@pytest.mark.parametrize("name,n", [("r02_bench_2gpu.json", 2), ("r02_bench_8gpu.json", 8)])
def test_multi_gpu_lines(name, n):
    d = _line(name)
    for k in BASE_KEYS:
        assert k in d, k
    assert d["n_gpus"] == n and d["scaling"] == "weak"
    assert abs(d["value"] - n * d["config"]["sizes"]["G"] / (d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"]
