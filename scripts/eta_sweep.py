"""KL_ORTHO_CGS2_SELECTIVE: iterations to tolerance, time and orthogonality over the threshold eta, against the
always-twice scheme (the reference's "Twice is enough", gmres_mgsr.f90:341) -- GMRES-MGSR(95) + cbpr2, rtol 1e-8.
Usage: python scripts/eta_sweep.py [ns ...]   (default 1024 2048)"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gmres_b200 as kl

P = (8.2, 0.2)
h = kl.Handle(0)
out = {}
for ns in [int(a) for a in sys.argv[1:]] or [1024, 2048]:
    b = h.apply(kl.stvec, torch.ones(ns * ns, dtype=torch.float64, device="cuda"), ns, ns)
    rows = {}
    for label, ortho, eta in (("cgs2 (always twice)", 1, 707), ("eta=0.707", 2, 707), ("eta=0.5", 2, 500), ("eta=0.3", 2, 300),
                              ("eta=0.1", 2, 100), ("eta=0.03", 2, 30), ("eta=0.01", 2, 10)):
        h.set_ortho(ortho)
        h.set_option(11, eta)
        h.set_option(3, 1)       # v_err epilogue: ||I - V^T V||_F of the last cycle
        t0 = time.perf_counter()
        r = h.gmres_mgsr_omp(kl.stvec, b, 95, 1e-8, kl.cbpr2, P, nx=ns, ny=ns)
        dt = time.perf_counter() - t0
        its = (r.restart_out - 1) * 95 + r.n_out
        rows[label] = dict(status=r.status, iterations=its, seconds=round(dt, 3), its_per_s=round(its / (r.stats["solve_ms"] * 1e-3), 1),
                           skipped=r.stats["reorth_skipped"], orth_frobenius=r.stats["orth_frobenius"],
                           linf_err=float((r.x - 1).abs().max()))
        print(ns, label, rows[label], flush=True)
    out[str(ns)] = rows
    del b
    torch.cuda.empty_cache()
h.set_ortho(1); h.set_option(11, 300)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/r2_eta_sweep.json", "w"), indent=1)
