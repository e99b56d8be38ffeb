#!/usr/bin/env python
"""Write gpurun_out/bench_history.json (to be committed as tests/golden/bench_history.json): the first residuals of
every bench workload from (a) the CPU oracle on this box's host cores and (b) the single-GPU CUDA path.  bench.py
compares the residual history of each timed run -- at any GPU count -- with these (config.parity)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
import gmres_b200 as kl
from oracle import oracle as ko

P = bench.P_REF
ko.build_native()
ko.set_threads(os.cpu_count())
h = kl.Handle(0)
out = {}
PLAN = {"cg16384": 64, "pcg16384": 64, "gmres4096": 95, "gmres4096sel": 95, "hh1024": 95, "gmres300": 475, "bicgstab8192": 20, "cg4096": 64}
only = sys.argv[1:] or list(PLAN)
for name in only:
    w = bench.WORKLOADS[name]
    k = PLAN[name]
    ns = w["nx"]
    A = ko.aniso_fn(*w["aniso"]) if w.get("aniso") else ko.stvec_fn()
    M = ko.cbpr2_fn() if w["pc"] else None
    b = ko.manufactured_rhs(A, ns)
    t0 = time.time()
    s = w["solver"]
    if s == "cg_omp":
        o = ko.cg_omp(A, b, 0.0, k)
    elif s == "pcg_omp":
        o = ko.pcg_omp(A, b, 0.0, k, M, P)
    elif s == "pbicgstab_omp":
        o = ko.pbicgstab_omp(A, b, 0.0, k, M, P)
    elif s == "gmres_mgsr_omp":
        o = ko.gmres_mgsr_omp(A, b, w["m"], 0.0, M, P, max_restarts=max(1, k // w["m"]), skip_verr=True)
    else:
        o = ko.gmres_hh(A, b, w["m"], 0.0, None, max_stages=max(1, k // w["m"]), skip_verr=True)
    t_cpu = time.time() - t0
    bt = torch.from_numpy(b).cuda()
    r = bench.run_gpu_solver(kl, h, w, bt, ns, ns, k)
    ho, hg = [float(v) for v in o.history[:k]], [float(v) for v in r.history[:k]]
    kk = min(len(ho), len(hg))
    rel = max(abs(hg[i] / ho[i] - 1) for i in range(kk))
    out[name] = {"oracle": ho, "gpu_n1": hg, "source": f"oracle/krylov_oracle.c ({os.cpu_count()} threads) and the single-GPU CUDA path, "
                 f"first {k} iterations, scripts/make_bench_history.py", "gpu_vs_oracle_max_rel": rel}
    print(f"{name}: {kk} iterations, cpu {t_cpu:.1f}s, gpu vs oracle max rel {rel:.2e}", flush=True)
    del bt
    torch.cuda.empty_cache()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "bench_history.json"), "w") as f:
    json.dump(out, f)
