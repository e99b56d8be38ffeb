"""Short workload for ncu: a few iterations of each fused solver at sizes >> L2."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import gmres_b200 as kl

which = sys.argv[1] if len(sys.argv) > 1 else "all"
h = kl.Handle(0)
h.set_option(3, 0)
def rhs(A, nx, ny):
    return h.apply(A, torch.ones(nx * ny, dtype=torch.float64, device="cuda"), nx, ny)
if which in ("all", "pcg"):
    n = 8192
    b = rhs(kl.stvec, n, n)
    h.set_option(4, 8)
    r = h.pcg_omp(kl.stvec, b, 0.0, 4, kl.cbpr2, (8.2, 0.2), nx=n, ny=n)
    print("pcg", r.stats["solve_ms"])
if which in ("all", "cg"):
    n = 8192
    b = rhs(kl.stvec, n, n)
    h.set_option(4, 8)
    r = h.cg_omp(kl.stvec, b, 0.0, 3, nx=n, ny=n)
    print("cg", r.stats["solve_ms"])
if which in ("all", "gmres"):
    n = 4096
    b = rhs(kl.stvec, n, n)
    h.set_option(2, 1)
    r = h.gmres_mgsr_omp(kl.stvec, b, 48, 0.0, kl.cbpr2, (8.2, 0.2), nx=n, ny=n)
    print("gmres", r.stats["solve_ms"])
if which in ("all", "bicg"):
    n = 8192
    b = rhs(kl.stvec, n, n)
    h.set_option(4, 8)
    r = h.pbicgstab_omp(kl.stvec, b, 0.0, 3, kl.cbpr2, (8.2, 0.2), nx=n, ny=n)
    print("bicg", r.stats["solve_ms"])
torch.cuda.synchronize()
