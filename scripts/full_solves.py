"""Full solves to tolerance at BASELINE.json's target sizes (north star: "GMRES-MGSR(m=95) on a 4096^2
Poisson grid and CG on a 16384^2 grid converge to the reference's answer").  x_true = 1, b = A*1."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gmres_b200 as kl

P = (8.2, 0.2)
h = kl.Handle(0)
h.set_option(3, 0)
out = {}
which = sys.argv[1:] or ["cg16384", "gmres4096"]
if "cg16384" in which:
    n = 16384
    b = h.apply(kl.stvec, torch.ones(n * n, dtype=torch.float64, device="cuda"), n, n)
    h.set_option(4, 256)
    t0 = time.perf_counter()
    r = h.cg_omp(kl.stvec, b, 1e-9, 100000, nx=n, ny=n)          # tests/test_cg.f90:20 tol (absolute)
    dt = time.perf_counter() - t0
    err = (r.x - 1).abs()
    out["cg_omp 16384^2 tol=1e-9 abs"] = dict(status=r.status, iterations=r.iter, res=r.res, seconds=dt,
                                               its_per_s=r.iter / (r.stats["solve_ms"] * 1e-3),
                                               linf_err=float(err.max()), l2_err=float(torch.linalg.vector_norm(r.x - 1)),
                                               roofline_frac=r.stats["algorithmic_bytes"] / (r.stats["solve_ms"] * 1e-3) / 1e9 / 6547.2)
    print(json.dumps(out), flush=True)
    del b, r, err
    torch.cuda.empty_cache()
if "gmres4096" in which:
    n, m = 4096, 95
    b = h.apply(kl.stvec, torch.ones(n * n, dtype=torch.float64, device="cuda"), n, n)
    lo, hi = h.lanczos(kl.stvec, n, n, 30)
    prm = h.cheb_params_from_ritz(lo, hi)
    t0 = time.perf_counter()
    r = h.gmres_mgsr_omp(kl.stvec, b, m, 1e-8, kl.cbpr2, P, nx=n, ny=n)
    dt = time.perf_counter() - t0
    its = (r.restart_out - 1) * m + r.n_out
    out["gmres_mgsr_omp(95)+cbpr2 4096^2 rtol=1e-8"] = dict(
        status=r.status, iterations=its, cycles=r.restart_out, n_out=r.n_out, final_err=float(r.final_err[r.n_out - 1]),
        seconds=dt, its_per_s=its / (r.stats["solve_ms"] * 1e-3), linf_err=float((r.x - 1).abs().max()),
        l2_err=float(torch.linalg.vector_norm(r.x - 1)), lanczos_ritz=(lo, hi), lanczos_params=prm,
        roofline_frac=r.stats["algorithmic_bytes"] / (r.stats["solve_ms"] * 1e-3) / 1e9 / 6547.2)
    print(json.dumps(out), flush=True)
for key in which:
    # degree-k Chebyshev preconditioner, k operator applications in one HBM pass (kl_chain_tma.cuh), interval
    # [b/ratio, b] with b = 1.025 * largest Lanczos(30) Ritz value:  gmres4096chebK / pcg16384chebK
    if "cheb" not in key:
        continue
    solver, k = key.split("cheb")[0], int(key.split("cheb")[1])
    n = 4096 if solver.startswith("gmres") else 16384
    m = 95
    b = h.apply(kl.stvec, torch.ones(n * n, dtype=torch.float64, device="cuda"), n, n)
    t0 = time.perf_counter()
    lo, hi = h.lanczos(kl.stvec, n, n, 30)
    prm = h.cheb_interval_from_ritz(hi, k)      # [b/ratio(k), b], b = 1.025*theta_max (ratio 1000 for k >= 5)
    if solver.startswith("gmres"):
        r = h.gmres_mgsr_omp(kl.stvec, b, m, 1e-8, kl.cheb(k), prm, nx=n, ny=n)
        its = (r.restart_out - 1) * m + r.n_out
        extra = dict(cycles=r.restart_out, n_out=r.n_out, final_err=float(r.final_err[r.n_out - 1]))
    else:
        h.set_option(4, 64)
        r = h.pcg_omp(kl.stvec, b, 1e-9, 200000, kl.cheb(k), prm, nx=n, ny=n)
        its = r.iter
        extra = dict(res=r.res)
    dt = time.perf_counter() - t0
    out[f"{solver}+cheb({k}) {n}^2 (Lanczos bounds, incl. the Lanczos run)"] = dict(
        status=r.status, iterations=its, seconds=dt, solve_seconds=r.stats["solve_ms"] * 1e-3,
        its_per_s=its / (r.stats["solve_ms"] * 1e-3), linf_err=float((r.x - 1).abs().max()),
        lanczos_ritz=(lo, hi), params=prm, **extra)
    print(json.dumps(out), flush=True)
    del b, r
    torch.cuda.empty_cache()
json.dump(out, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "full_solves.json"), "w"), indent=1)
