"""Kernel-only CG timing on one GPU at the slab size a rank owns in the 8-GPU strong-scaling run
(16384 x 2048) for several CTA heights (KL_OPT_STENCIL_ROWS)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gmres_b200 as kl

nx = 16384
h = kl.Handle(0)
h.set_option(3, 0)
for ny in (2048, 4096, 16384):
    b = h.apply(kl.stvec, torch.ones(nx * ny, dtype=torch.float64, device="cuda"), nx, ny)
    for rows in (0, 32, 64):
        h.set_option(kl.KL_OPT_STENCIL_ROWS, rows)
        h.set_option(4, 50)
        h.cg_omp(kl.stvec, b, 0.0, 10, nx=nx, ny=ny)
        h.set_option(8, 1)
        h.cg_omp(kl.stvec, b, 0.0, 50, nx=nx, ny=ny)
        h.set_option(8, 0)
        prof = {p["name"].split(" ")[0]: 1e3 * p["ms"] / p["launches"] for p in h.profile()}
        r = h.cg_omp(kl.stvec, b, 0.0, 50, nx=nx, ny=ny)
        us = r.stats["solve_ms"] * 1e3 / 50
        ideal = 64.0 * nx * ny / 6547.2e9 * 1e6
        print(f"{nx}x{ny} rows {rows:4d}: {us:7.1f} us/it ({100 * ideal / us:5.1f}% of roofline)  " +
              "  ".join(f"{k} {v:6.1f}us" for k, v in prof.items()), flush=True)
    h.set_option(kl.KL_OPT_STENCIL_ROWS, 0)
    del b
h.close()
