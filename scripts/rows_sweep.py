"""CTA-height sweep of the temporally blocked kernels (KL_OPT_STENCIL_ROWS): BiCGSTAB+cbpr2 8192^2 and cheb(k)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gmres_b200 as kl

n = 8192
st = torch.cuda.Stream()
h = kl.Handle(0, stream=st.cuda_stream)
h.set_option(3, 0)
with torch.cuda.stream(st):
    A = kl.aniso(1.0, 0.01)
    b = h.apply(A, torch.ones(n * n, dtype=torch.float64, device="cuda"), n, n)
    z = torch.empty_like(b)
    for rows in (0, 36, 64, 96, 128, 192, 256):
        h.set_option(kl.KL_OPT_STENCIL_ROWS, rows)
        h.set_option(4, 30)
        h.pbicgstab_omp(A, b, 0.0, 5, kl.cbpr2, (8.2, 0.2), nx=n, ny=n)
        h.set_option(8, 1)
        r = h.pbicgstab_omp(A, b, 0.0, 30, kl.cbpr2, (8.2, 0.2), nx=n, ny=n)
        h.set_option(8, 0)
        prof = {p["name"].split(" ")[0]: 1e3 * p["ms"] / p["launches"] for p in h.profile()}
        r = h.pbicgstab_omp(A, b, 0.0, 30, kl.cbpr2, (8.2, 0.2), nx=n, ny=n)
        line = f"rows {rows:4d}: bicgstab {30 / (r.stats['solve_ms'] * 1e-3):7.1f} it/s  " + "  ".join(f"{k[:12]} {v:6.1f}us" for k, v in prof.items())
        for k in (2, 4):
            for _ in range(2):
                h.set_output_buffer(z); h.apply_precond(kl.cheb(k), kl.stvec, b, (0.2, 8.2), n, n)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            for _ in range(10):
                h.set_output_buffer(z); h.apply_precond(kl.cheb(k), kl.stvec, b, (0.2, 8.2), n, n)
            e1.record(st); st.synchronize()
            line += f"  cheb({k}) {e0.elapsed_time(e1) * 100:6.1f}us"
        print(line, flush=True)
h.close()
