"""Multi-GPU parity check (run under torchrun, one rank per GPU): the row-slab solve over
NCCL must match the single-GPU solve of the same global problem."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import gmres_b200 as kl
from gmres_b200.dist import init_handle, slab_partition

P = (8.2, 0.2)
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
h = init_handle(local)
h1 = kl.Handle(local)          # single-GPU reference on every rank (same device)
ok = True
if rank == 0:
    print("peer-memory collectives:", bool(h.get_option(10)), flush=True)

def gather(xl, nx, ny):
    parts = [None] * world
    dist.all_gather_object(parts, xl)
    return np.concatenate(parts)

def check(name, cond, extra=""):
    global ok
    ok = ok and bool(cond)
    if rank == 0:
        print(f"[{'ok' if cond else 'FAIL'}] {name} {extra}", flush=True)

for (nx, ny) in ((512, 512), (300, 301), (1000, 64)):
    j0, nyl = h.partition(ny)
    assert (j0, nyl) == slab_partition(ny, rank, world)
    rng = np.random.default_rng(3)
    xg = rng.standard_normal(nx * ny)
    xl = xg.reshape(ny, nx)[j0:j0 + nyl].reshape(-1).copy()
    for A in (kl.stvec, kl.aniso(1.0, 0.01)):
        y = gather(h.apply(A, xl, nx, ny), nx, ny)
        check(f"apply {nx}x{ny} kind={A.kind}", np.array_equal(y, h1.apply(A, xg, nx, ny)))
    z = gather(h.apply_precond(kl.cbpr2, kl.stvec, xl, P, nx, ny), nx, ny)
    check(f"cbpr2 {nx}x{ny}", np.array_equal(z, h1.apply_precond(kl.cbpr2, kl.stvec, xg, P, nx, ny)))
    for k in (2, 3, 6, 7):   # temporally blocked Chebyshev with a k-line halo (k <= 6 in one pass)
        zc = gather(h.apply_precond(kl.cheb(k), kl.stvec, xl, (0.2, 8.2), nx, ny), nx, ny)
        check(f"cheb({k}) chain {nx}x{ny}", np.array_equal(zc, h1.apply_precond(kl.cheb(k), kl.stvec, xg, (0.2, 8.2), nx, ny)))
    bg = h1.apply(kl.stvec, np.ones(nx * ny), nx, ny)
    bl = bg.reshape(ny, nx)[j0:j0 + nyl].reshape(-1).copy()
    for name, run in (
        ("cg_omp", lambda hh, b: hh.cg_omp(kl.stvec, b, 1e-9, 20000, nx=nx, ny=ny)),
        ("pcg_omp", lambda hh, b: hh.pcg_omp(kl.stvec, b, 1e-9, 20000, kl.cbpr2, P, nx=nx, ny=ny)),
        ("pbicgstab_omp", lambda hh, b: hh.pbicgstab_omp(kl.stvec, b, 1e-9, 20000, kl.cbpr2, P, nx=nx, ny=ny)),
        ("bicgstab", lambda hh, b: hh.bicgstab(kl.stvec, b, 1e-9, 20000, nx=nx, ny=ny)),
    ):
        m, s = run(h, bl), run(h1, bg)
        x = gather(m.x, nx, ny)
        tol_it = 1 if "cg" in name and "bi" not in name else max(3, int(0.2 * s.iter))
        check(f"{name} {nx}x{ny}", m.status == 0 and abs(m.iter - s.iter) <= tol_it and np.abs(x - 1).max() < 1e-7,
              f"iters multi {m.iter} single {s.iter} dx {np.abs(x - s.x).max():.2e}")
    for ortho in (1, 0):
        h.set_ortho(ortho); h1.set_ortho(ortho)
        m = h.gmres_mgsr_omp(kl.stvec, bl, 40, 1e-8, kl.cbpr2, P, nx=nx, ny=ny)
        s = h1.gmres_mgsr_omp(kl.stvec, bg, 40, 1e-8, kl.cbpr2, P, nx=nx, ny=ny)
        x = gather(m.x, nx, ny)
        mi, si = (m.restart_out - 1) * 40 + m.n_out, (s.restart_out - 1) * 40 + s.n_out
        k = min(m.history.size, s.history.size)
        check(f"gmres_mgsr_omp ortho={ortho} {nx}x{ny}", m.status == 0 and abs(mi - si) <= 1 and
              np.abs(x - s.x).max() < 1e-9 and np.abs(m.history[:k] / s.history[:k] - 1).max() < 1e-7,
              f"iters multi {mi} single {si} dx {np.abs(x - s.x).max():.2e} verr {m.v_err[m.n_out]:.2e}")
    h.set_ortho(1); h1.set_ortho(1)
    lo, hi = h.lanczos(kl.stvec, nx, ny, 20)
    lo1, hi1 = h1.lanczos(kl.stvec, nx, ny, 20)
    check(f"lanczos {nx}x{ny}", abs(hi - hi1) < 1e-9 and abs(lo - lo1) < 1e-7, f"{lo} {hi} vs {lo1} {hi1}")
dist.barrier()
if rank == 0:
    print("MGPU_CHECK", "PASS" if ok else "FAIL", flush=True)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
