"""Kernel-only CG timing on one GPU at slab sizes of the strong-scaling run for CTA heights (KL_OPT_STENCIL_ROWS)
and tapered-tail heights (KL_OPT_STENCIL_TAIL)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gmres_b200 as kl

nx = 16384
h = kl.Handle(0)
h.set_option(3, 0)
import json
SWEEP = json.loads(os.environ["KL_SWEEP"]) if os.environ.get("KL_SWEEP") else {
    "2048": [(0, -1), (0, 0), (32, 8), (32, 0), (32, 16), (34, 0), (34, 8), (40, 10), (48, 12), (64, 16), (64, 8), (24, 0), (24, 8), (16, 0)],
    "4096": [(0, -1), (0, 0), (64, 16), (48, 12)],
    "16384": [(0, -1), (0, 0), (64, 16), (128, 32), (133, 0)]}
for ny, combos in ((int(k), v) for k, v in SWEEP.items()):
    b = h.apply(kl.stvec, torch.ones(nx * ny, dtype=torch.float64, device="cuda"), nx, ny)
    torch.cuda.synchronize()
    for rows, tail in combos:
        h.set_option(13, rows)
        h.set_option(17, tail)
        h.set_option(4, 50)
        h.cg_omp(kl.stvec, b, 0.0, 10, nx=nx, ny=ny)
        best = 1e9
        for rep in range(3):
            r = h.cg_omp(kl.stvec, b, 0.0, 50, nx=nx, ny=ny)
            best = min(best, r.stats["solve_ms"] * 1e3 / 50)
        h.set_option(8, 1)
        h.cg_omp(kl.stvec, b, 0.0, 50, nx=nx, ny=ny)
        h.set_option(8, 0)
        prof = {p["name"].split(" ")[0]: 1e3 * p["ms"] / p["launches"] for p in h.profile()}
        ideal = 64.0 * nx * ny / 6547.2e9 * 1e6
        print(f"{nx}x{ny} rows {rows:4d} tail {tail:3d}: {best:7.1f} us/it ({100 * ideal / best:5.1f}% of roofline)  " +
              "  ".join(f"{k} {v:6.1f}us" for k, v in prof.items()), flush=True)
    h.set_option(13, 0); h.set_option(17, -1)
    del b
h.close()
