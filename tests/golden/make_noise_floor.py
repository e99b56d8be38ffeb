#!/usr/bin/env python
"""Measure the reference's own reproducibility floor on the CPU oracle and write tests/golden/noise_floor.json.

The reference's OpenMP reductions (`!$omp do reduction(+:...)`, cg.f90:118-133, gmres_mgsr.f90:346-351,
bicgstab.f90:123-127) sum in a thread-count-dependent order, so two runs of the same reference binary with
different OMP_NUM_THREADS give different residual histories.  The oracle keeps that loop / reduction structure,
so oracle(1 thread) vs oracle(T threads) IS that floor.  The GPU parity tests use max(1e-10, 2 x floor).
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as ko  # noqa: E402

P = (8.2, 0.2)


def rel(a, b, floor=1e-12):
    k = min(a.size, b.size)
    a, b = a[:k], b[:k]
    m = np.abs(b) > floor
    return float(np.max(np.abs(a[m] / b[m] - 1.0))) if m.any() else 0.0


def rel_first(a, b, k):
    return rel(a[:k], b[:k])


out = {"threads": [1, max(2, min(ko.max_threads(), 8))], "note": "max relative difference of the residual history, "
       "oracle with 1 thread vs T threads; *_first50 = first 50 iterations / first restart cycle"}
T = out["threads"][1]
for ns in (128, 300):
    b = ko.manufactured_rhs(ko.stvec_fn(), ns)
    runs = {}
    for t in (1, T):
        ko.set_threads(t)
        runs[t] = dict(
            cg=ko.cg_omp(ko.stvec_fn(), b, 1e-9, 100000),
            pcg=ko.pcg_omp(ko.stvec_fn(), b, 1e-9, 100000, ko.cbpr2_fn(), P),
            gm=ko.gmres_mgsr_omp(ko.stvec_fn(), b, 95, 1e-8, ko.cbpr2_fn(), P),
            hh=ko.gmres_hh(ko.stvec_fn(), b, 95, 1e-8, ko.cbpr2_fn(), P),
            bi=ko.pbicgstab_omp(ko.stvec_fn(), b, 1e-9, 100000, ko.cbpr2_fn(), P))
    ko.set_threads(1)
    a, c = runs[1], runs[T]
    for key, name in (("cg", "cg_omp"), ("pcg", "pcg_omp"), ("gm", "gmres_mgsr_omp"), ("hh", "gmres_hh_prec_omp"),
                      ("bi", "pbicgstab_omp")):
        ra, rc = a[key], c[key]
        ia = ra.iter if hasattr(ra, "iter") else ra.iterations
        ic = rc.iter if hasattr(rc, "iter") else rc.iterations
        out[f"{name}_{ns}"] = dict(iterations=[int(ia), int(ic)], history_rel=rel(rc.history, ra.history),
                                   history_rel_first50=rel_first(rc.history, ra.history, 50),
                                   x_maxdiff=float(np.max(np.abs(ra.x - rc.x))))
        print(name, ns, out[f"{name}_{ns}"], flush=True)
with open(os.path.join(ROOT, "tests", "golden", "noise_floor.json"), "w") as f:
    json.dump(out, f, indent=1)
