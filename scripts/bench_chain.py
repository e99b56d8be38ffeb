"""Micro-benchmark of the temporally blocked Chebyshev kernel (kl_chain_tma.cuh) against the
one-pass-per-application kernels: time of z = Cheb_k(A) r on device-resident vectors, CUDA events on
the handle's stream.  Usage: python scripts/bench_chain.py [ns] [reps]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import gmres_b200 as kl

ns = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
peak = 6547.2
try:
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
st = torch.cuda.Stream()
h = kl.Handle(0, stream=st.cuda_stream)
n = ns * ns
with torch.cuda.stream(st):
    r = torch.randn(n, dtype=torch.float64, device="cuda")
    z = torch.empty_like(r)
    out = {}
    for k in (1, 2, 3, 4, 5, 6, 8, 12):
        for chain in (1, 0):
            h.set_option(kl.KL_OPT_CHAIN, chain)
            h.set_output_buffer(z)
            for _ in range(3):
                h.set_output_buffer(z)
                h.apply_precond(kl.cheb(k), kl.stvec, r, (0.2, 8.2), ns, ns)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            for _ in range(reps):
                h.set_output_buffer(z)
                h.apply_precond(kl.cheb(k), kl.stvec, r, (0.2, 8.2), ns, ns)
            e1.record(st)
            st.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / reps
            gbs = 16.0 * n / us / 1e3
            out[f"k{k}_{'chain' if chain else 'stepwise'}"] = dict(us=round(us, 1), gbs_16n=round(gbs, 1),
                                                                  frac_16n=round(gbs / peak, 3))
        a, b = out[f"k{k}_chain"]["us"], out[f"k{k}_stepwise"]["us"]
        print(f"cheb({k}) {ns}^2: chain {a:9.1f} us ({out[f'k{k}_chain']['frac_16n']*100:5.1f}% of the 16n roofline)"
              f"   stepwise {b:9.1f} us   speed-up {b / a:5.2f}x", flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(dict(ns=ns, reps=reps, peak_gbs=peak, results=out), open("gpurun_out/bench_chain.json", "w"), indent=1)
h.close()
