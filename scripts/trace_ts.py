"""Debug: per-CTA time line of the last tall-skinny pass (needs the -DKL_TRACE build,
KRYLOV_B200_LIB=.../libkrylov_b200_trace.so).  Usage: python scripts/trace_ts.py ns m [ns m ...]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import gmres_b200 as kl
from gmres_b200.api import load_library

L = load_library()
L.kl_debug_trace_ts.argtypes = [C.c_void_p, C.c_int]
h = kl.Handle(0)
h.set_option(5, 0)      # no graph replay: plain launches
args = [int(a) for a in sys.argv[1:]] or [300, 95, 300, 48, 300, 12, 1024, 95]
for ns, m in zip(args[0::2], args[1::2]):
    b = h.apply(kl.stvec, torch.ones(ns * ns, dtype=torch.float64, device="cuda"), ns, ns)
    h.set_option(2, 2)
    for _ in range(2):
        r = h.gmres_mgsr_omp(kl.stvec, b, m, 0.0, kl.cbpr2, (8.2, 0.2), nx=ns, ny=ns)
    buf = np.zeros(8 * 512, dtype=np.uint64)
    assert L.kl_debug_trace_ts(buf.ctypes.data, buf.size) == 0
    t = buf.reshape(-1, 8).astype(np.int64)
    t = t[t[:, 0] > 0]
    t0 = t[:, 0].min()
    rel = (t[:, :5] - t0) / 1e3
    last = t[:, 4] > t0
    print(f"{ns}^2 m={m}: {len(t)} CTAs, us/step {r.stats['solve_ms'] * 1e3 / r.stats['iterations']:.1f}")
    for k, name in enumerate(("entry", "first tile landed", "tile loop done", "arrived at counter")):
        print(f"   {name:20s} median {np.median(rel[:, k]):6.2f}  p95 {np.percentile(rel[:, k], 95):6.2f}  max {rel[:, k].max():6.2f} us")
    if last.any():
        print(f"   last block: sums written at {rel[last, 4].max():6.2f} us")
h.close()
