"""Debug: per-CTA time line of the fused CG kernels (needs the -DKL_TRACE build, KRYLOV_B200_LIB=.../libkrylov_b200_trace.so).
Prints when CTAs start and end relative to the first start, per 'wave', and the idle tail."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import gmres_b200 as kl
from gmres_b200.api import load_library

nx, ny = 16384, int(sys.argv[1]) if len(sys.argv) > 1 else 2048
L = load_library()
h = kl.Handle(0)
h.set_option(3, 0)
b = h.apply(kl.stvec, torch.ones(nx * ny, dtype=torch.float64, device="cuda"), nx, ny)
torch.cuda.synchronize()
h.set_option(4, 50)
h.cg_omp(kl.stvec, b, 0.0, 10, nx=nx, ny=ny)
r = h.cg_omp(kl.stvec, b, 0.0, 20, nx=nx, ny=ny)
print("us/it", r.stats["solve_ms"] * 1e3 / 20)
buf = np.zeros(3 * 16384 + 8, dtype=np.uint64)
L.kl_debug_trace_cg.argtypes = [C.c_void_p, C.c_int]
assert L.kl_debug_trace_cg(buf.ctypes.data, buf.size) == 0
t = buf[:3 * 16384].reshape(-1, 3)
nb = int((t[:, 0] > 0).sum())
t = t[:nb]
st, en = t[:, 0].astype(np.int64), t[:, 1].astype(np.int64)
sm = (t[:, 2] >> np.uint64(48)).astype(np.int64)
wait = (t[:, 2] & np.uint64((1 << 48) - 1)).astype(np.int64)
t0 = st.min()
st, en = (st - t0) / 1e3, (en - t0) / 1e3
tail_end = (int(buf[3 * 16384]) - t0) / 1e3
print(f"CTAs {nb}  SMs {len(set(sm))}  first start 0  last start {st.max():.1f} us  last end {en.max():.1f} us  last-block tail end {tail_end:.1f} us")
print(f"CTA lifetime us: min {np.min(en - st):.1f} median {np.median(en - st):.1f} max {np.max(en - st):.1f} ; griddep wait median {np.median(wait) / 1e3:.2f} max {wait.max() / 1e3:.2f}")
# concurrency over time
edges = np.linspace(0, en.max(), 41)
for a, bb in zip(edges[:-1], edges[1:]):
    mid = 0.5 * (a + bb)
    active = int(((st <= mid) & (en > mid)).sum())
    print(f"  t={mid:7.1f} us  active CTAs {active:5d}  started {int((st <= mid).sum()):5d}  done {int((en <= mid).sum()):5d}")
