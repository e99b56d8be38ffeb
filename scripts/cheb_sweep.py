"""Degree / interval sweep of the temporally blocked Chebyshev preconditioner: PCG and GMRES(95) iterations
and time-to-tolerance on a Poisson grid (x_true = 1, b = A*1).  Usage: python scripts/cheb_sweep.py [ns]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gmres_b200 as kl

ns = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
h = kl.Handle(0)
h.set_option(3, 0)
h.set_option(4, 64)
b = h.apply(kl.stvec, torch.ones(ns * ns, dtype=torch.float64, device="cuda"), ns, ns)
lo, hi = h.lanczos(kl.stvec, ns, ns, 30)
top = 1.025 * hi
out = {"ns": ns, "ritz": (lo, hi), "rows": []}
print(f"grid {ns}^2, Lanczos(30) Ritz values [{lo:.4g}, {hi:.5g}]", flush=True)
for solver in ("pcg", "gmres"):
    for k, ratios in ((0, (41,)), (1, (41,)), (2, (41, 100)), (4, (41, 100, 400)), (6, (100, 400, 1000))):
        for ratio in ratios:
            M = kl.cbpr2 if k == 0 else kl.cheb(k)
            prm = (top, top / ratio)
            torch.cuda.synchronize()
            if solver == "pcg":
                r = h.pcg_omp(kl.stvec, b, 1e-9, 200000, M, prm, nx=ns, ny=ns)
                its, ok = r.iter, r.status == 0
            else:
                r = h.gmres_mgsr_omp(kl.stvec, b, 95, 1e-8, M, prm, nx=ns, ny=ns)
                its, ok = (r.restart_out - 1) * 95 + r.n_out, r.status == 0
            ms = r.stats["solve_ms"]
            err = float((r.x - 1).abs().max())
            row = dict(solver=solver, precond="cbpr2" if k == 0 else f"cheb({k})", ratio=ratio, iterations=its,
                       ok=ok, solve_ms=ms, linf_err=err)
            out["rows"].append(row)
            print(f"{solver:5s} {row['precond']:8s} [b/{ratio:<4d}, b]  its {its:7d}  {ms:9.1f} ms  "
                  f"{ms / max(its, 1) * 1e3:7.1f} us/it  err {err:.2e} {'' if ok else 'NOT CONVERGED'}", flush=True)
json.dump(out, open("gpurun_out/cheb_sweep.json", "w"), indent=1)
