"""Parity metrics shared by the CPU and GPU tests.

The north star asks for "residual history within 1e-10 relative".  Two readings are asserted:

  hist_norm(a, b, r0) = max_k |a_k - b_k| / ||r_0||      over the whole history; bar 1e-10.
      (GMRES's final_err is already relative to beta0 = ||b||, gmres_mgsr.f90:383, so r0 = 1 there.)
  hist_rel(a, b)      = max_k |a_k / b_k - 1|            point-wise.
      bar 1e-10 over the first restart cycle / first 50 iterations;
      bar max(1e-10, 2 x floor) over the whole history, where `floor` is what the REFERENCE ITSELF shows
      between a 1-thread and a T-thread run (tests/golden/noise_floor.json, measured on the oracle, which has the
      reference's OpenMP reduction structure): late in a solve the point-wise ratio of ANY two correct
      implementations grows as the residual falls towards the rounding level of the recurrences.
"""
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def hist_rel(a, b, floor=1e-12):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    k = min(a.size, b.size)
    a, b = a[:k], b[:k]
    m = np.abs(b) > floor
    return float(np.max(np.abs(a[m] / b[m] - 1.0))) if m.any() else 0.0


def hist_norm(a, b, r0=1.0):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    k = min(a.size, b.size)
    return float(np.max(np.abs(a[:k] - b[:k]))) / r0 if k else 0.0


_floor = None


def noise_floor(key, field="history_rel"):
    """the reference's own 1-thread vs T-thread drift for a case, e.g. 'cg_omp_300', 'gmres_mgsr_omp_300_95'"""
    global _floor
    if _floor is None:
        with open(os.path.join(ROOT, "tests", "golden", "noise_floor.json")) as f:
            _floor = json.load(f)
    return float(_floor[key][field])


def hist_bar(key):
    """whole-history point-wise bar for the CUDA path: max(1e-10, 2 x the reference's own floor)"""
    return max(1e-10, 2.0 * noise_floor(key))


def x_diff(x, c):
    """max difference between a computed solution and a stored golden one (full vector, or head + sums)"""
    x = np.asarray(x)
    if "x" in c:
        return float(np.max(np.abs(x - np.array(c["x"]))))
    h = np.array(c["x_head"])
    d = float(np.max(np.abs(x[: h.size] - h)))
    d = max(d, abs(float(np.sum(x)) - c["x_sum"]) / x.size)
    return max(d, abs(float(np.max(np.abs(x - 1.0))) - c["x_err_inf"]))
