"""The CUDA path (through the C ABI) against OUTPUT OF THE REFERENCE ITSELF.

tests/golden/reference_f90.json = what the reference's own Fortran sources return when executed through
oracle/f90run (tests/golden/make_reference_golden.py).  Same inputs as the reference's drivers (x = 1, b = A*1,
params (8.2, 0.2)); bars: identical iteration counts above the rounding floor (+-1 at tol 1e-15), residual
history within 1e-10 (metrics in tests/parity.py), solution within 1e-9 relative -- the north star's numbers.
"""
import json
import os

import numpy as np
import pytest

from parity import hist_norm, hist_rel, x_diff

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = (8.2, 0.2)


@pytest.fixture(scope="module")
def ref():
    with open(os.path.join(ROOT, "tests", "golden", "reference_f90.json")) as f:
        return json.load(f)["cases"]


@pytest.fixture(scope="module")
def kl():
    import gmres_b200 as m
    return m


@pytest.fixture(scope="module")
def h(kl):
    hd = kl.Handle(0)
    yield hd
    hd.close()


def cases(ref, prefix):
    out = [(k, v) for k, v in ref.items() if k.startswith(prefix + "_") and k[len(prefix) + 1][0].isdigit()]
    assert out, prefix
    return out


def rhs(kl, h, ns):
    return h.apply(kl.stvec, np.ones(ns * ns), ns, ns)      # call stvec(x, b, nsize)  test_poisson_mf.f90:39-40


def test_operators_match_the_reference(kl, h, ref):
    for key, c in cases(ref, "operators"):
        ns, x = c["ns"], np.array(c["x"])
        assert np.array_equal(h.apply(kl.stvec, x, ns, ns), np.array(c["stvec"])), key
        assert np.array_equal(h.apply(kl.stv_poisson, x, ns, ns), np.array(c["stv_poisson"])), key
        z = h.apply_precond(kl.cbpr2, kl.stvec, x, P, ns, ns)
        assert np.max(np.abs(z - np.array(c["cbpr2"]))) < 2.3e-16 * np.max(np.abs(x)), key   # FMA vs mul+add


@pytest.mark.parametrize("ortho", [0, 1])
@pytest.mark.parametrize("name", ["gmres_mgsr_omp", "gmres_mgsr_mf"])
def test_gmres_mgsr_matches_the_reference(kl, h, ref, name, ortho):
    h.set_ortho(ortho)
    try:
        for key, c in cases(ref, name):
            ns, m, tol = c["ns"], c["m"], c["tol"]
            b = rhs(kl, h, ns)
            g = getattr(h, name)(kl.stvec, b, m, tol, kl.cbpr2, P)
            its = (g.restart_out - 1) * m + g.n_out
            assert g.status == 0, key
            if tol >= 1e-12:
                assert (its, g.n_out, g.restart_out) == (c["iterations"], c["n_out"], c["stages"]), key
            else:
                assert abs(its - c["iterations"]) <= 1, key
            assert x_diff(g.x, c) < 1e-9, key
            if "history" in c and tol >= 1e-12:
                k1 = min(m, 50)
                assert hist_rel(g.history[:k1], c["history"][:k1], 1e-9) < 1e-10, (key, ortho)
                assert hist_norm(g.history, c["history"]) < 1e-10, (key, ortho)
    finally:
        h.set_ortho(1)


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("name", ["gmres_hh_prec_omp", "gmres_hh_omp"])
def test_gmres_hh_matches_the_reference(kl, h, ref, name, mode):
    h.set_option(6, mode)       # KL_OPT_HH_MODE: sequential reflectors / compact-WY
    try:
        for key, c in cases(ref, name):
            ns, m, tol = c["ns"], c["m"], c["tol"]
            b = rhs(kl, h, ns)
            g = h.gmres_hh_prec_omp(kl.stvec, b, m, tol, kl.cbpr2, P) if name.endswith("prec_omp") else \
                h.gmres_hh_omp(kl.stvec, b, m, tol)
            its = (g.restart_out - 1) * m + g.n_out
            unprec = name == "gmres_hh_omp"
            if tol >= 1e-12:
                # gmres_hh_omp only tests at cycle ends (gmres_hh.f90:340-344 commented out): whole cycles
                assert abs(its - c["iterations"]) <= (m if unprec else 0), key
            else:
                assert abs(its - c["iterations"]) <= (m if unprec else 1), key
            assert x_diff(g.x, c) < (1e-7 if unprec else 1e-9), key
            if "history" in c and tol >= 1e-12:
                k1 = min(m, 50)
                assert hist_rel(g.history[:k1], c["history"][:k1], 1e-9) < 1e-10, (key, mode)
                if not unprec:
                    assert hist_norm(g.history, c["history"]) < 1e-10, (key, mode)
            assert g.v_err.max() < 1e-27, key     # README.md:10, calculate_verr's squared metric
    finally:
        h.set_option(6, 1)


@pytest.mark.parametrize("name", ["cg", "cg_omp", "pcg", "pcg_omp"])
def test_cg_matches_the_reference(kl, h, ref, name):
    for key, c in cases(ref, name):
        ns = c["ns"]
        b = rhs(kl, h, ns)
        fn = getattr(h, name)
        g = fn(kl.stvec, b, c["tol"], 100000, kl.cbpr2, P) if name.startswith("p") else fn(kl.stvec, b, c["tol"], 100000)
        assert g.status == 0 and abs(g.iter - c["iterations"]) <= 1, (key, g.iter, c["iterations"])
        assert x_diff(g.x, c) < 1e-9, key
        if "history" in c:
            assert hist_rel(g.history[:50], c["history"][:50], 1e-9) < 1e-10, key
            assert hist_norm(g.history, c["history"], float(np.linalg.norm(b))) < 1e-10, key


@pytest.mark.parametrize("name", ["bicgstab", "pbicgstab", "pbicgstab_omp"])
def test_bicgstab_matches_the_reference(kl, h, ref, name):
    for key, c in cases(ref, name):
        ns = c["ns"]
        b = rhs(kl, h, ns)
        fn = getattr(h, name)
        g = fn(kl.stvec, b, c["tol"], 100000, kl.cbpr2, P) if name.startswith("p") else fn(kl.stvec, b, c["tol"], 100000)
        # the reference's own count moves by several per cent with its thread count (noise_floor.json)
        assert g.status == 0 and abs(g.iter - c["iterations"]) <= max(2, c["iterations"] // 10), key
        assert x_diff(g.x, c) < 1e-7, key
        if "history" in c:
            k = min(8, len(c["history"]), g.history.size)
            assert hist_rel(g.history[:k], c["history"][:k]) < 1e-9, key


def test_dense_variants_match_the_reference(kl, h, ref):
    for key, c in ref.items():
        if not (key.startswith("gmres_mgsr_dense_poisson") or key.startswith("gmres_hh_dense_poisson")):
            continue
        ns, m = c["ns"], c["m"]
        n = ns * ns
        A = np.zeros((n, n))
        for col in range(n):
            e = np.zeros(n)
            e[col] = 1.0
            A[:, col] = h.apply(kl.stvec, e, ns, ns)
        b = A @ np.ones(n)
        g = (h.gmres_mgsr_dense if "mgsr" in key else h.gmres_hh_dense)(A, b, m, c["tol"])
        assert ((g.restart_out - 1) * m + g.n_out) == c["iterations"], key
        assert x_diff(g.x, c) < 1e-10, key
    for n in (4, 8):
        Hm = np.array(ref[f"hilbert_{n}"]["H"]).reshape(n, n, order="F")
        assert np.array_equal(h.generate_matrix(n), Hm)


def test_driver_program_numbers(kl, h, ref):
    """what tests/test_poisson_mf.f90 prints (executed as a program): iterations, stages, L_max error."""
    for key, c in ref.items():
        if not key.startswith("program_test_poisson_mf"):
            continue
        ns, m = int(c["argv"][0]), int(c["argv"][1])
        recs = [r for r in c["records"] if r]
        its = [r for r in recs if r[0] == "Iterations until convergence:"]
        lmax = [r for r in recs if r[0] == "Max error L_max:"]
        b = rhs(kl, h, ns)
        hh = h.gmres_hh_prec_omp(kl.stvec, b, m, 1e-15, kl.cbpr2, P)       # test_poisson_mf.f90:45
        mg = h.gmres_mgsr_omp(kl.stvec, b, m, 1e-15, kl.cbpr2, P)          # :76
        for g, r_it, r_l, ck in ((hh, its[0], lmax[0], "hh_prec_cycle1_final_err"), (mg, its[1], lmax[1], "mgsr_cycle1_final_err")):
            g_it = (g.restart_out - 1) * m + g.n_out
            if r_it[3] == 1 or ck not in c:
                assert abs(g_it - r_it[1]) <= 1, key
            else:
                # several cycles at tol 1e-15 (the 100 x 95 run): the first cycle is the reproducible part (1e-10
                # against the reference's own first-cycle history); the later cycles start from a residual at the
                # rounding level of x and their length is not reproducible even between two runs of the reference
                # (see the fixture's note and tests/test_reference_golden.py)
                assert abs(g_it - r_it[1]) <= 0.2 * r_it[1], (key, g_it, r_it)
                assert hist_rel(g.history[:m], np.array(c[ck]), 1e-9) < 1e-10, (key, ck)
            assert abs(g.restart_out - r_it[3]) <= 1, key
            assert np.max(np.abs(g.x - 1.0)) < 10 * max(r_l[1], 1e-14), key
