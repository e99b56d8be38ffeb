#!/usr/bin/env python
"""Measure the reference's own reproducibility floor on the CPU oracle and write tests/golden/noise_floor.json.

The reference's OpenMP reductions (`!$omp do reduction(+:...)`, cg.f90:118-133, gmres_mgsr.f90:346-351,
gmres_hh.f90:455-459, bicgstab.f90:123-127) sum in a thread-count-dependent order, so two runs of the same
reference binary with different OMP_NUM_THREADS give different residual histories.  The oracle keeps that
loop / reduction structure, so oracle(1 thread) vs oracle(T threads) IS that floor.  The GPU parity tests hold
the CUDA path to max(1e-10, 2 x floor) on the point-wise relative history difference (tests/parity.py).
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import oracle as ko  # noqa: E402
from parity import hist_rel, hist_norm  # noqa: E402

P = (8.2, 0.2)
T = max(2, min(ko.max_threads(), 8))
REPS = 6
out = {"threads": [1, T], "reps": REPS, "note": "oracle with 1 thread vs T and T/2 threads, worst of REPS runs: history_rel = max |h_k/h'_k - 1| over the whole "
       "history (residuals above 1e-12), history_rel_head = the same over the first 50 iterations / first restart "
       "cycle, history_norm = max |h_k - h'_k| / ||r_0||"}

CASES = [("cg_omp", ns, 0) for ns in (100, 128, 300)] + [("pcg_omp", ns, 0) for ns in (100, 128, 300)] + \
        [("gmres_mgsr_omp", 100, 95), ("gmres_mgsr_omp", 128, 95), ("gmres_mgsr_omp", 300, 95), ("gmres_mgsr_omp", 300, 50),
         ("gmres_hh_prec_omp", 100, 95), ("gmres_hh_prec_omp", 100, 20), ("gmres_hh_prec_omp", 128, 24),
         ("gmres_hh_prec_omp", 128, 95), ("gmres_hh_prec_omp", 300, 95),
         ("pbicgstab_omp", 100, 0), ("pbicgstab_omp", 300, 0)]


def run(name, b, m):
    A, M = ko.stvec_fn(), ko.cbpr2_fn()
    if name == "cg_omp":
        return ko.cg_omp(A, b, 1e-9, 100000)
    if name == "pcg_omp":
        return ko.pcg_omp(A, b, 1e-9, 100000, M, P)
    if name == "gmres_mgsr_omp":
        return ko.gmres_mgsr_omp(A, b, m, 1e-8, M, P)
    if name == "gmres_hh_prec_omp":
        return ko.gmres_hh(A, b, m, 1e-8, M, P)
    return ko.pbicgstab_omp(A, b, 1e-9, 100000, M, P)


for name, ns, m in CASES:
    b = ko.manufactured_rhs(ko.stvec_fn(), ns)
    ko.set_threads(1)
    a = run(name, b, m)
    ia = a.iter if hasattr(a, "iter") else a.iterations
    head = min(m, 50) if m else 50
    r0 = 1.0 if name.startswith("gmres") else float(np.linalg.norm(b))
    key = f"{name}_{ns}" + (f"_{m}" if m else "")
    rec = dict(iterations=[int(ia)], history_rel=0.0, history_rel_head=0.0, history_norm=0.0, x_maxdiff=0.0)
    # the order in which OpenMP combines the threads' partial sums varies from run to run: worst of REPS runs
    for rep in range(REPS):
        ko.set_threads(T if rep % 2 == 0 else max(2, T // 2))
        c = run(name, b, m)
        rec["iterations"].append(int(c.iter if hasattr(c, "iter") else c.iterations))
        rec["history_rel"] = max(rec["history_rel"], hist_rel(c.history, a.history))
        rec["history_rel_head"] = max(rec["history_rel_head"], hist_rel(c.history[:head], a.history[:head]))
        rec["history_norm"] = max(rec["history_norm"], hist_norm(c.history, a.history, r0))
        rec["x_maxdiff"] = max(rec["x_maxdiff"], float(np.max(np.abs(a.x - c.x))))
    ko.set_threads(1)
    out[key] = rec
    print(key, out[key], flush=True)
with open(os.path.join(ROOT, "tests", "golden", "noise_floor.json"), "w") as f:
    json.dump(out, f, indent=1)
