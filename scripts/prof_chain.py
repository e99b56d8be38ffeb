"""Short workload for ncu: the temporally blocked kernels and the CG kernels with the deferred x update."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gmres_b200 as kl

h = kl.Handle(0)
h.set_option(3, 0)
n = 8192
ones = torch.ones(n * n, dtype=torch.float64, device="cuda")
b = h.apply(kl.stvec, ones, n, n)
h.set_option(4, 8)
r = h.pbicgstab_omp(kl.stvec, b, 0.0, 2, kl.cbpr2, (8.2, 0.2), nx=n, ny=n)
print("bicg", r.stats["solve_ms"])
r = h.cg_omp(kl.stvec, b, 0.0, 2, nx=n, ny=n)
print("cg", r.stats["solve_ms"])
r = h.pcg_omp(kl.stvec, b, 0.0, 2, kl.cbpr2, (8.2, 0.2), nx=n, ny=n)
print("pcg", r.stats["solve_ms"])
z = torch.empty_like(b)
for k in (2, 4):
    h.set_output_buffer(z)
    h.apply_precond(kl.cheb(k), kl.stvec, b, (0.2, 8.2), n, n)
torch.cuda.synchronize()
