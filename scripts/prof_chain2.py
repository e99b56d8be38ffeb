"""Workload for ncu: the degree-k Chebyshev chain kernel alone, three launches at ns^2.
Usage: python scripts/prof_chain2.py k [ns]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import gmres_b200 as kl

k = int(sys.argv[1])
ns = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
h = kl.Handle(0)
r = torch.randn(ns * ns, dtype=torch.float64, device="cuda")
z = torch.empty_like(r)
for _ in range(3):
    h.set_output_buffer(z)
    h.apply_precond(kl.cheb(k), kl.stvec, r, (0.2, 8.2), ns, ns)
torch.cuda.synchronize()
print("ok", float(z.abs().max()))
