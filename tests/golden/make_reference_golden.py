#!/usr/bin/env python
"""Generate tests/golden/reference_f90.json by EXECUTING THE REFERENCE'S OWN FORTRAN SOURCES.

The reference (AlexanderGSC/gmres) is Fortran-only and no Fortran compiler exists in the build container or
on the GPU box.  oracle/f90run translates the reference's source files mechanically (statement by statement,
no algorithmic knowledge) into Python and runs them with one thread; this script drives the translated module
procedures and driver programs on the reference's own manufactured problem (x = 1, b = A*1,
tests/test_poisson_mf.f90:39-40; params = (8.2, 0.2), :38) and stores what they return.

This is the only file that reads /root/reference; the JSON it writes is committed and is what the tests use
(the GPU box has no /root/reference).  Run:   python tests/golden/make_reference_golden.py [--quick]

Residual histories: the reference returns final_err of the LAST restart cycle only (gmres_mgsr.f90:302,383) and
`res` of the last CG/BiCGSTAB iteration, so per-cycle / per-iteration histories are obtained by re-running with
the restart cap (module variables max_restarts / stages, gmres_mgsr.f90:6, gmres_hh.f90:8) or the iteration
cap (`iter` on entry, cg.f90:15) set to 1, 2, 3, ...
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = os.environ.get("KRYLOV_REFERENCE", "/root/reference")
PARAMS = (8.2, 0.2)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true", help="small cases only (seconds)")
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden", "reference_f90.json"))
    args = ap.parse_args()

    from oracle.f90run import f90py
    w = f90py.load_reference(REF, with_tests=True)

    def call(module, name, **kw):
        """call a translated module procedure by dummy-argument names; returns {dummy: value} of its outputs"""
        sub = w.units[module].subs[name]
        actual = []
        for a in sub.args:
            v = kw.get(a)
            dv = sub.vars[a]
            if v is None and dv.rank == 0 and dv.typ in ("integer",):
                v = 0
            elif v is None and dv.rank == 0 and dv.typ == "real8":
                v = np.float64(0.0)
            elif isinstance(v, float):
                v = np.float64(v)
            actual.append(v)
        ret = w.proc(module, name)(*actual)
        return {sub.args[k]: ret[j] for j, k in enumerate(sub.out_positions())}

    stvec = w.proc("poisson", "stvec")
    stv_poisson = w.proc("poisson", "stv_poisson")
    cbpr2 = w.proc("chebyshev_precond", "cbpr2")
    params = np.array(PARAMS)

    def rhs(ns):
        x, b = np.ones(ns * ns), np.zeros(ns * ns)
        stvec(x, b, ns)
        return b

    def lst(a):
        return [float(v) for v in np.asarray(a).reshape(-1, order="F")]

    def xrec(x, ns):
        """the solution: in full for small grids, as head + sums for the larger ones (keeps the fixture small)"""
        x = np.asarray(x)
        if ns <= 48:
            return dict(x=lst(x))
        return dict(x_head=lst(x[:64]), x_sum=float(np.sum(x)), x_err_inf=float(np.max(np.abs(x - 1.0))),
                    x_err_l2=float(np.sqrt(np.sum((x - 1.0) ** 2))))

    G = {"generator": "tests/golden/make_reference_golden.py", "reference": "AlexanderGSC/gmres (Fortran sources executed "
         "through oracle/f90run, one thread, no FMA contraction)", "params": list(PARAMS), "cases": {}}
    C = G["cases"]
    t_all = time.time()

    # ---- operators and preconditioner on a seeded vector ------------------------------------------------
    for ns in (5, 13, 24):
        rng = np.random.default_rng(ns)
        x = rng.standard_normal(ns * ns)
        y1, y2, z, aux = np.zeros_like(x), np.zeros_like(x), np.zeros_like(x), np.zeros_like(x)
        stvec(x, y1, ns)
        stv_poisson(x, y2, ns)
        cbpr2(stvec, x, z, aux, params, ns)
        C[f"operators_{ns}"] = dict(ns=ns, seed=ns, x=lst(x), stvec=lst(y1), stv_poisson=lst(y2), cbpr2=lst(z))

    # ---- GMRES variants -------------------------------------------------------------------------------------
    def gmres_case(module, name, ns, m, tol, capvar, prec, cycles_hist=True):
        b = rhs(ns)
        kw = dict(b=b, m=m, tol=tol)
        opname = "ax_vec"
        kw[opname] = stvec
        if prec:
            kw["m_inv"] = cbpr2
            kw["params"] = params
        nsm = w.ns[module]
        default_cap = nsm[capvar]
        t0 = time.time()
        r = call(module, name, **kw)
        stages_key = "restart_out" if "restart_out" in r else "stages_out"
        rec = dict(ns=ns, m=m, tol=tol, n_out=int(r["n_out"]), stages=int(r[stages_key]),
                   iterations=(int(r[stages_key]) - 1) * m + int(r["n_out"]),
                   final_err=lst(r["final_err"]), v_err=lst(r["v_err"]), **xrec(r["x"], ns))
        if cycles_hist:
            hist = []
            for k in range(1, rec["stages"] + 1):
                nsm[capvar] = k
                rk = call(module, name, **kw)
                n_k = m if k < rec["stages"] else rec["n_out"]
                hist.extend(lst(rk["final_err"])[:n_k])
            nsm[capvar] = default_cap
            rec["history"] = hist
        rec["seconds"] = round(time.time() - t0, 2)
        return rec

    gm = [(16, 10, 1e-8), (16, 10, 1e-15), (24, 20, 1e-8), (32, 30, 1e-10)]
    if not args.quick:
        gm += [(48, 30, 1e-8), (64, 95, 1e-8), (100, 95, 1e-8)]      # 100^2: the preconditioned variants only (below)
    for ns, m, tol in gm:
        tag = f"{ns}_{m}_{tol:g}"
        C[f"gmres_mgsr_omp_{tag}"] = gmres_case("gmres_mgsr_mod", "gmres_mgsr_omp", ns, m, tol, "max_restarts", True)
        C[f"gmres_mgsr_mf_{tag}"] = gmres_case("gmres_mgsr_mod", "gmres_mgsr_mf", ns, m, tol, "max_restarts", True,
                                               cycles_hist=ns <= 24)
        C[f"gmres_hh_prec_omp_{tag}"] = gmres_case("gmres_hh_mod", "gmres_hh_prec_omp", ns, m, tol, "stages", True,
                                                   cycles_hist=ns <= 32)
        if ns <= 32:      # unpreconditioned Householder GMRES needs hundreds of cycles on the larger grids
            C[f"gmres_hh_omp_{tag}"] = gmres_case("gmres_hh_mod", "gmres_hh_omp", ns, m, tol, "stages", False,
                                                  cycles_hist=ns <= 24)
        print(f"gmres {tag}: {C[f'gmres_mgsr_omp_{tag}']['iterations']} its "
              f"(hh_prec {C[f'gmres_hh_prec_omp_{tag}']['iterations']})  {time.time() - t_all:.0f}s", flush=True)

    # ---- CG / BiCGSTAB ------------------------------------------------------------------------------------------
    def cg_case(module, name, ns, tol, prec, opkey, itkey, hist):
        b = rhs(ns)
        kw = {opkey: stvec, "b": b, "tol": tol, itkey: 100000}
        if prec:
            kw["m_inv"] = cbpr2
            kw["params"] = params
        t0 = time.time()
        r = call(module, name, **kw)
        its = int(r[itkey])
        rec = dict(ns=ns, tol=tol, iterations=its, res=float(r["res"]), **xrec(r["x"], ns))
        if hist:
            h = []
            for k in range(1, its + 1):
                kw[itkey] = k
                h.append(float(call(module, name, **kw)["res"]))
            rec["history"] = h
        rec["seconds"] = round(time.time() - t0, 2)
        return rec

    cgs = [(16, True), (32, True)] + ([] if args.quick else [(64, False), (100, False)])
    for ns, hist in cgs:
        for name, prec in (("cg", False), ("cg_omp", False), ("pcg", True), ("pcg_omp", True)):
            C[f"{name}_{ns}"] = cg_case("conjugate_gradient", name, ns, 1e-9, prec, "ax_op", "iter", hist and name.endswith("omp"))
        C[f"bicgstab_{ns}"] = cg_case("bicgstab_mod", "bicgstab", ns, 1e-9, False, "ax_op", "iter", False)
        C[f"pbicgstab_{ns}"] = cg_case("bicgstab_mod", "pbicgstab", ns, 1e-9, True, "ax_op", "iter", False)
        C[f"pbicgstab_omp_{ns}"] = cg_case("bicgstab_mod", "pbicgstab_omp", ns, 1e-9, True, "ax_op", "max_iter", hist)
        print(f"cg {ns}: cg_omp {C[f'cg_omp_{ns}']['iterations']} pcg_omp {C[f'pcg_omp_{ns}']['iterations']} "
              f"pbicgstab_omp {C[f'pbicgstab_omp_{ns}']['iterations']}  {time.time() - t_all:.0f}s", flush=True)

    # ---- the reference drivers' own grid (tests/test_cg.f90:21, tests/test_bicgstab.f90: nsize = 300), OpenMP variants only
    # (70-90 s each in the interpreter; no per-iteration history: that would be one run per iteration)
    if not args.quick:
        C["pcg_omp_300"] = cg_case("conjugate_gradient", "pcg_omp", 300, 1e-9, True, "ax_op", "iter", False)
        C["cg_omp_300"] = cg_case("conjugate_gradient", "cg_omp", 300, 1e-9, False, "ax_op", "iter", False)
        C["pbicgstab_omp_300"] = cg_case("bicgstab_mod", "pbicgstab_omp", 300, 1e-9, True, "ax_op", "max_iter", False)
        print(f"300^2: pcg_omp {C['pcg_omp_300']['iterations']} cg_omp {C['cg_omp_300']['iterations']} "
              f"pbicgstab_omp {C['pbicgstab_omp_300']['iterations']}  {time.time() - t_all:.0f}s", flush=True)

    # ---- dense variants (dense 5-point matrix and Hilbert) ---------------------------------------------------------
    for ns, m in ((6, 10), (8, 20)):
        A = call("poisson", "generate_matrix", nsize=ns)["a"]
        b = A @ np.ones(ns * ns)
        for module, name, key in (("gmres_mgsr_mod", "gmres_mgsr_dense", "restart_out"), ("gmres_hh_mod", "gmres_hh_dense", "stages_out")):
            r = call(module, name, a=A, b=b, m=m, tol=1e-10)
            C[f"{name}_poisson_{ns}_{m}"] = dict(ns=ns, m=m, tol=1e-10, n_out=int(r["n_out"]), stages=int(r[key]),
                                                 iterations=(int(r[key]) - 1) * m + int(r["n_out"]), x=lst(r["x"]),
                                                 final_err=lst(r["final_err"]), v_err=lst(r["v_err"]))
    for n in (4, 8):
        Hm = call("hilbert", "generate_matrix", n=n)["h"]
        C[f"hilbert_{n}"] = dict(n=n, H=lst(Hm))

    # ---- the reference's own driver program, tests/test_poisson_mf.f90 (argv = grid size, restart length) -------------
    # ("100", "95"): the reference's own restart length and tolerance (README.md:20 runs 300 95; 300^2 takes hours in the
    # interpreter, 100^2 takes 100 s)
    for argv in (("16", "10"), ("24", "20")) + (() if args.quick else (("100", "95"),)):
        out = w.run_program("test_poisson_mf", argv)
        rec = dict(
            argv=list(argv), records=[r["items"] for r in out if r["items"] and not str(r["items"][0]).startswith("Elapsed")])
        if argv == ("100", "95"):
            # the second cycle of a tol = 1e-15 solve starts from a residual at the rounding level of x and is not
            # reproducible (see the note): keep the first-cycle histories of both solvers as the pinned quantity
            ns_, m_ = int(argv[0]), int(argv[1])
            for module, name, capvar, key in (("gmres_hh_mod", "gmres_hh_prec_omp", "stages", "hh_prec_cycle1_final_err"),
                                              ("gmres_mgsr_mod", "gmres_mgsr_omp", "max_restarts", "mgsr_cycle1_final_err")):
                default_cap = w.ns[module][capvar]
                w.ns[module][capvar] = 1
                r1 = call(module, name, ax_vec=stvec, b=rhs(ns_), m=m_, tol=1e-15, m_inv=cbpr2, params=params)
                w.ns[module][capvar] = default_cap
                rec[key] = lst(r1["final_err"])
            rec["note"] = ("first-cycle histories (restart cap 1) of gmres_hh_prec_omp / gmres_mgsr_omp at the driver's tol 1e-15; "
                           "the second cycle starts from a residual at the rounding level of x (1.3e-9 relative), so its length "
                           "is not reproducible (reference 54 iterations, the C oracle 72 from an initial residual that differs "
                           "by 6e-8 relative)")
        C["program_test_poisson_mf_" + "_".join(argv)] = rec

    G["seconds_total"] = round(time.time() - t_all, 1)
    with open(args.out, "w") as f:
        json.dump(G, f, separators=(",", ":"))
    print("wrote", args.out, os.path.getsize(args.out), "bytes in", G["seconds_total"], "s")


if __name__ == "__main__":
    main()
