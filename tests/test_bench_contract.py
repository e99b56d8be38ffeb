"""The JSON line bench.py prints must keep the driver's contract.  Checked on the committed round-1 and round-2
lines (profiles/): no GPU needed."""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(name):
    with open(os.path.join(ROOT, "profiles", name)) as f:
        for ln in f:
            if ln.startswith("{"):
                return json.loads(ln)
    raise AssertionError(name)


@pytest.mark.parametrize("name,n", [("r01_bench_n1_final.json", 1), ("r01_bench_cg16384_n2_64n.json", 2),
                                    ("r01_bench_cg16384_n4_64n.json", 4), ("r01_bench_cg16384_n8_64n.json", 8),
                                    ("r02_bench_n1.json", 1), ("r02_bench_cg16384_n2.json", 2),
                                    ("r02_bench_cg16384_n4.json", 4), ("r02_bench_cg16384_n8.json", 8)])
def test_bench_line_contract(name, n):
    d = _line(name)
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline"):
        assert k in d, k
    assert d["n_gpus"] == n and d["unit"] == "iterations/s" and d["dtype"] == "f64" and d["higher_is_better"] is True
    assert d["vs_baseline"] is None and d["scaling"] == "strong" and "workload" in d["config"]
    assert d["gpu_launches"] > 0 and d["steps"] >= 1 and d["warmup"] >= 3
    assert abs(d["value"] - 1e3 / d["ms_per_step"]) < 1e-6 * d["value"]
    e = d["e2e"]
    assert e["unit"] == d["unit"] and 0 < e["value"] < d["value"]            # host buffers + copies: slower than resident
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert 0.5 < r["frac"] < 1.1
    c = d["clocks"]
    assert c["sm_mhz"] and c["sm_max_mhz"] and not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown",
                                                                        "sw_thermal_slowdown"}
    if name.startswith("r02"):
        # round 2: every line, at every N, carries the parity of the timed run against the committed oracle history,
        # and the N = 1 line carries the other workloads
        par = d["config"]["parity"]
        assert par["iterations_compared"] >= 20 and par["max_rel"] < 1e-10, par
        if n == 1:
            ex = d["config"]["extras"]
            assert {"gmres4096", "gmres4096sel", "pcg16384", "bicgstab8192", "hh1024", "gmres300"} <= set(ex)
            assert all(v["its_per_s"] > 0 for v in ex.values())
    if n == 1:
        # DRAM traffic measured by ncu agrees with the algorithmic bytes of the dominant kernel
        assert r["traffic"] and abs(r["traffic"] / r["algorithmic_bytes_per_launch"] - 1) < 0.05
        b = d["cpu_baseline"]
        assert b["kind"] == "port" and b["cores"] >= 1 and b["value"] > 0 and b["sample"]
        assert d["value"] > 20 * b["value"]


def test_strong_scaling_meets_the_north_star():
    one = _line("r01_bench_n1_final.json")["value"]
    eight = _line("r01_bench_cg16384_n8_64n.json")["value"]
    assert eight / one >= 6.0          # BASELINE.json: ">= 6x strong-scaling speedup at 8 GPUs for the CG case"
    # round 2, all four N on ONE box
    v = {n: _line(f"r02_bench_cg16384_n{n}.json")["value"] for n in (1, 2, 4, 8)}
    assert v[8] / v[1] >= 6.8 and v[4] / v[1] >= 3.6 and v[2] / v[1] >= 1.9
