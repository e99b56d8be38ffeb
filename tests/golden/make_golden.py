#!/usr/bin/env python
"""Generate tests/golden/*.json|npz from an INDEPENDENT numpy restatement.

The reference (Fortran) cannot be built or imported in the build container
(no gfortran), and it ships no golden vectors, so the fixtures here come from a
second, separately written restatement of the same reference routines in
vectorised numpy (BLAS/pairwise summation => different rounding than the C
oracle's sequential sums).  Agreement between the two restatements (iteration
counts equal, residuals/solutions equal to ~1e-8 relative) plus the analytic
facts the reference's drivers rely on (x == 1, ||b|| = sqrt(4 n + 8)) is what
pins the oracle.  Run:  python tests/golden/make_golden.py

Reference lines followed are cited per function.
"""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def stvec(x, n):
    """poisson.f90:33-77 (zero-padded form; same neighbour order l,r,+n,-n)."""
    X = x.reshape(n, n)  # X[j, i], i fastest
    P = np.zeros((n + 2, n + 2))
    P[1:-1, 1:-1] = X
    s = ((P[1:-1, :-2] + P[1:-1, 2:]) + P[2:, 1:-1]) + P[:-2, 1:-1]
    return (4.0 * X - s).reshape(-1)


def cbpr2(A, r, params, n):
    """chebyshev.f90:8-38."""
    emin, emax = params
    c = (emax - emin) / 2.0
    d = (emax + emin) / 2.0
    alpha = 1.0 / d
    beta = (c * alpha / 2.0) ** 2
    alpha = 1.0 / (d - beta)
    z = r / d
    aux = A(z, n)
    return z + alpha * (r - aux)


def givens(H, cs, sn, g, j):
    """gmres_mgsr.f90:365-380."""
    for i in range(j):
        tmp = H[i, j]
        H[i, j] = cs[i] * tmp + sn[i] * H[i + 1, j]
        H[i + 1, j] = -sn[i] * tmp + cs[i] * H[i + 1, j]
    ds = np.hypot(H[j + 1, j], H[j, j])
    cs[j] = H[j, j] / ds
    sn[j] = H[j + 1, j] / ds
    H[j, j] = cs[j] * H[j, j] + sn[j] * H[j + 1, j]
    H[j + 1, j] = 0.0
    tmp = g[j]
    g[j] = cs[j] * tmp + sn[j] * g[j + 1]
    g[j + 1] = -sn[j] * tmp + cs[j] * g[j + 1]


def backsolve(H, g, n_out):
    y = np.zeros(n_out)
    y[n_out - 1] = g[n_out - 1] / H[n_out - 1, n_out - 1]
    for i in range(n_out - 2, -1, -1):
        y[i] = (g[i] - H[i, i + 1:n_out] @ y[i + 1:n_out]) / H[i, i]
    return y


def gmres_mgsr_omp(A, b, m, tol, M, params, max_restarts=1000):
    """gmres_mgsr.f90:277-421."""
    n = b.size
    ns = int(np.sqrt(np.float32(n)))
    x = np.zeros(n)
    final_err = np.zeros(m)
    beta0 = np.linalg.norm(b)
    hist = []
    converged = False
    restart_out = max_restarts
    for st in range(1, max_restarts + 1):
        V = np.zeros((m + 1, n))
        H = np.zeros((m + 1, m))
        g = np.zeros(m + 1)
        cs = np.zeros(m)
        sn = np.zeros(m)
        w = M(A, b - A(x, ns), params, ns)
        beta = np.linalg.norm(w)
        g[0] = beta
        V[0] = w / beta
        for j in range(m):
            if converged:
                continue
            w = M(A, A(V[j], ns), params, ns)
            for _ in range(2):
                for i in range(j + 1):
                    h = w @ V[i]
                    H[i, j] += h
                    w = w - h * V[i]
            h_val = np.linalg.norm(w)
            H[j + 1, j] = h_val
            givens(H, cs, sn, g, j)
            final_err[j] = abs(g[j + 1]) / beta0
            hist.append(final_err[j])
            V[j + 1] = w / h_val
            if final_err[j] < tol:
                restart_out = st
                converged = True
            n_out = j + 1
        y = backsolve(H, g, n_out)
        x = x + V[:n_out].T @ y
        if h_val < tol or final_err[n_out - 1] < tol:
            restart_out = st
            break
    return dict(x=x, n_out=n_out, restart_out=restart_out, history=np.array(hist))


def gmres_hh(A, b, m, tol, M, params, max_stages=1000):
    """gmres_hh.f90:211-385 (M None) / :388-566 (M given)."""
    n = b.size
    ns = int(np.sqrt(np.float32(n)))
    x = np.zeros(n)
    final_err = np.zeros(m)
    beta0 = np.linalg.norm(b)
    hist = []
    converged = False
    for k in range(1, max_stages + 1):
        P = np.zeros((m + 1, n))
        H = np.zeros((m + 1, m))
        g = np.zeros(m + 1)
        cs = np.zeros(m)
        sn = np.zeros(m)
        w = b - A(x, ns)
        if M is not None:
            w = M(A, w, params, ns)
        beta = np.linalg.norm(w)
        g[0] = -np.copysign(beta, w[0])
        w[0] = np.copysign(beta, w[0]) + w[0]
        P[0] = w / np.linalg.norm(w)
        for j in range(m):
            if M is not None and converged:
                continue
            n_out = j + 1
            v = np.zeros(n)
            v[j] = 1.0
            for i in range(j, -1, -1):
                v = v - 2.0 * P[i] * (v @ P[i])
            w = A(v, ns)
            if M is not None:
                w = M(A, w, params, ns)
            for i in range(j + 1):
                w = w - 2.0 * P[i] * (w @ P[i])
            H[:j + 1, j] = w[:j + 1]
            tmp = np.linalg.norm(w[j + 1:])
            H[j + 1, j] = -tmp if w[j + 1] > 0.0 else tmp
            w[:j + 1] = 0.0
            w[j + 1] = w[j + 1] - H[j + 1, j]
            w = w / np.linalg.norm(w)
            P[j + 1] = w
            givens(H, cs, sn, g, j)
            final_err[j] = abs(g[j + 1]) / beta0
            hist.append(final_err[j])
            if M is not None and final_err[j] < tol:
                converged = True
        y = backsolve(H, g, n_out)
        w = np.zeros(n)
        w[:n_out] = y
        for i in range(n_out - 1, -1, -1):
            w = w - 2.0 * P[i] * (P[i] @ w)
        x = x + w
        stages_out = k
        if final_err[n_out - 1] < tol:
            break
    # calculate_verr, gmres_hh.f90:568-593
    V = np.zeros((n_out, n))
    for i in range(n_out):
        V[i, i] = 1.0
        for j in range(i, -1, -1):
            V[i] = V[i] - 2.0 * P[j] * (V[i] @ P[j])
    v_err = np.zeros(m + 1)
    G = V @ V.T
    for i in range(1, n_out):
        v_err[i] = np.sum(2.0 * G[i, :i] ** 2)
    return dict(x=x, n_out=n_out, stages_out=stages_out, history=np.array(hist),
                v_err_max=float(v_err.max()))


def cg_omp(A, b, tol, max_iter, M=None, params=None):
    """cg.f90:83-152 (M None) / :154-234."""
    n = b.size
    ns = int(np.sqrt(np.float32(n)))
    x = np.zeros(n)
    r = b.copy()
    z = M(A, r, params, ns) if M is not None else r
    p = z.copy()
    hist = []
    it = max_iter
    for i in range(1, max_iter + 1):
        ax = A(p, ns)
        rr = r @ z
        alpha = rr / (ax @ p)
        x = x + alpha * p
        r = r - alpha * ax
        res = np.sqrt(r @ r)
        z = M(A, r, params, ns) if M is not None else r
        beta = (r @ z) / rr
        p = z + beta * p
        hist.append(res)
        if res < tol:
            it = i
            break
    return dict(x=x, iter=it, res=res, history=np.array(hist))


def pbicgstab_omp(A, b, tol, max_iter, M=None, params=None):
    """bicgstab.f90:91-182."""
    n = b.size
    ns = int(np.sqrt(np.float32(n)))
    x = np.zeros(n)
    r = b.copy()
    r0 = r.copy()
    p = r.copy()
    hist = []
    it = max_iter
    for i in range(1, max_iter + 1):
        z1 = M(A, p, params, ns) if M is not None else p
        ap = A(z1, ns)
        rr0 = r @ r0
        alpha = rr0 / (ap @ r0)
        s = r - alpha * ap
        z2 = M(A, s, params, ns) if M is not None else s
        as_ = A(z2, ns)
        omega = (as_ @ s) / (as_ @ as_)
        x = x + alpha * z1 + omega * z2
        r = s - omega * as_
        res = np.linalg.norm(r)
        hist.append(res)
        if res < tol:
            it = i
            break
        beta = ((r @ r0) / rr0) * (alpha / omega)
        p = r + beta * (p - omega * ap)
    return dict(x=x, iter=it, res=res, history=np.array(hist))


def main():
    params = (8.2, 0.2)  # tests/test_poisson_mf.f90:38
    out = {"params": params, "cases": {}}
    rng = np.random.default_rng(0)
    # --- operator / preconditioner vectors (bit-level fixtures) ---
    n = 37
    xv = rng.standard_normal(n * n)
    np.savez_compressed(os.path.join(HERE, "stencil_37.npz"), x=xv, y_stvec=stvec(xv, n),
                        z_cbpr2=cbpr2(stvec, xv, params, n))
    for ns in (100, 300):
        b = stvec(np.ones(ns * ns), ns)
        c = {"norm_b": float(np.linalg.norm(b)), "norm_b_exact": float(np.sqrt(4 * ns + 8))}
        for tol in (1e-8,):
            g = gmres_mgsr_omp(stvec, b, 95, tol, cbpr2, params)
            c[f"gmres_mgsr_omp_m95_tol{tol:g}"] = dict(
                iterations=(g["restart_out"] - 1) * 95 + g["n_out"], n_out=g["n_out"],
                restart_out=g["restart_out"], final_err=float(g["history"][-1]),
                linf=float(np.abs(g["x"] - 1).max()), l2=float(np.linalg.norm(g["x"] - 1)),
                history_head=[float(v) for v in g["history"][:60]])
            h = gmres_hh(stvec, b, 95, tol, cbpr2, params)
            c[f"gmres_hh_prec_omp_m95_tol{tol:g}"] = dict(
                iterations=(h["stages_out"] - 1) * 95 + h["n_out"], n_out=h["n_out"],
                stages_out=h["stages_out"], final_err=float(h["history"][-1]),
                linf=float(np.abs(h["x"] - 1).max()), v_err_max=h["v_err_max"],
                history_head=[float(v) for v in h["history"][:60]])
        if ns == 100:
            h = gmres_hh(stvec, b, 30, 1e-8, None, None, max_stages=3)
            c["gmres_hh_omp_m30_3stages"] = dict(
                iterations=(h["stages_out"] - 1) * 30 + h["n_out"], final_err=float(h["history"][-1]),
                v_err_max=h["v_err_max"], history_head=[float(v) for v in h["history"][:60]])
        r = cg_omp(stvec, b, 1e-9, 10000)
        c["cg_omp_tol1e-9"] = dict(iter=r["iter"], res=float(r["res"]),
                                   linf=float(np.abs(r["x"] - 1).max()),
                                   history_head=[float(v) for v in r["history"][:60]])
        r = cg_omp(stvec, b, 1e-9, 10000, cbpr2, params)
        c["pcg_omp_tol1e-9"] = dict(iter=r["iter"], res=float(r["res"]),
                                    linf=float(np.abs(r["x"] - 1).max()),
                                    history_head=[float(v) for v in r["history"][:60]])
        r = pbicgstab_omp(stvec, b, 1e-9, 10000, cbpr2, params)
        c["pbicgstab_omp_tol1e-9"] = dict(iter=r["iter"], res=float(r["res"]),
                                          linf=float(np.abs(r["x"] - 1).max()),
                                          history_head=[float(v) for v in r["history"][:30]])
        out["cases"][str(ns)] = c
        print(ns, json.dumps({k: (v if not isinstance(v, dict) else {kk: vv for kk, vv in v.items() if kk != "history_head"}) for k, v in c.items()}, indent=1))
    with open(os.path.join(HERE, "kat_numpy.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
