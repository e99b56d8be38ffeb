"""Isolate which option changes a result: GMRES-MGSR+cbpr2 and Householder GMRES under option toggles."""
import os, sys, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import gmres_b200 as kl
from oracle import oracle as ko

P = (8.2, 0.2)
ko.set_threads(8)
h = kl.Handle(0)
OPT = dict(chain=12, graph=5, coop=20, persistent=21)
for nx, ny, m in ((1024, 256, 24), (512, 512, 24)):
    b = h.apply(kl.stvec, np.ones(nx * ny), nx, ny)
    base = None
    for chain, graph, coop, pers in itertools.product((1, 0), (1, 0), (1, 0), (1, 0)):
        for k, v in zip(("chain", "graph", "coop", "persistent"), (chain, graph, coop, pers)):
            h.set_option(OPT[k], v)
        h.set_option(2, 40)
        r = h.gmres_mgsr_omp(kl.stvec, b, m, 1e-9, kl.cbpr2, P, nx=nx, ny=ny)
        its = (r.restart_out - 1) * m + r.n_out
        if base is None and (chain, graph, coop, pers) == (0, 0, 0, 0):
            pass
        print(f"gmres {nx}x{ny} chain={chain} graph={graph} coop={coop} persistent={pers}: status {r.status} its {its} "
              f"hist[:3] {r.history[:3]} hist[{m}] {r.history[min(m, r.history.size - 1)]:.6e} err {np.abs(r.x - 1).max():.2e}", flush=True)
for k in OPT.values():
    h.set_option(k, 1)
h.set_option(2, 1000)
ns, m = 256, 24
b = ko.manufactured_rhs(ko.stvec_fn(), ns)
o = ko.gmres_hh(ko.stvec_fn(), b, m, 1e-8, None)
for mode, graph, pers in itertools.product((1, 0), (1, 0), (1, 0)):
    h.set_option(6, mode); h.set_option(5, graph); h.set_option(21, pers)
    r = h.gmres_hh_omp(kl.stvec, b, m, 1e-8)
    k = min(r.history.size, o.history.size, 3 * m)
    print(f"hh {ns} mode={mode} graph={graph} persistent={pers}: status {r.status} its {(r.restart_out - 1) * m + r.n_out} (oracle {o.iterations}) "
          f"hist rel {np.abs(r.history[:k] / o.history[:k] - 1).max():.2e} x diff {np.abs(r.x - o.x).max():.2e}", flush=True)
