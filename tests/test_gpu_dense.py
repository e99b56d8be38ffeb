"""GPU tests of the dense-operator variants (gmres_mgsr_dense, gmres_hh_dense, generate_matrix):
src/gmres_mgsr.f90:11-95, src/gmres_hh.f90:10-112, src/problems/hilbert.f90:6-18, tests/test_hilbert.f90."""
import numpy as np
import pytest

from test_oracle import _dense_poisson

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def kl():
    import gmres_b200 as m
    return m


@pytest.fixture(scope="module")
def h(kl):
    hd = kl.Handle(0)
    yield hd
    hd.close()


def test_generate_matrix_and_matvec_bit_exact(kl, h, ko):
    for n in (5, 12, 100, 257):
        H = h.generate_matrix(n)
        assert np.array_equal(H, ko.generate_matrix(n))
        x = np.random.default_rng(n).standard_normal(n)
        assert np.array_equal(h.dense_matvec(H, x), ko.dense_matvec(H, x))


@pytest.mark.parametrize("ns,m", [(12, 30), (20, 40)])
def test_dense_poisson_parity(kl, h, ko, ns, m):
    A = _dense_poisson(ns)
    b = ko.manufactured_rhs(ko.stvec_fn(), ns)
    assert np.array_equal(h.dense_matvec(A, np.ones(ns * ns)), b)
    h.set_ortho(kl.ORTHO_MGS2)
    try:
        g = h.gmres_mgsr_dense(A, b, m, 1e-10)
    finally:
        h.set_ortho(kl.ORTHO_CGS2)
    o = ko.gmres_mgsr_dense(A, b, m, 1e-10)
    gi, oi = (g.restart_out - 1) * m + g.n_out, o.iterations
    k = min(g.history.size, o.history.size)
    print(f"gmres_mgsr_dense {ns}^2: gpu {gi} oracle {oi}")
    assert g.status == 0 and abs(gi - oi) <= 1
    big = o.history[:k] > 1e-8         # entries near rounding level (the last ones) are noise
    assert np.abs(g.history[:k][big] / o.history[:k][big] - 1).max() < 1e-6
    assert np.abs(g.x - o.x).max() < 1e-9 and np.abs(g.x - 1).max() < 1e-8
    # the dense solver is the matrix-free one with matmul as the operator: same counts as the stencil path
    f = h.gmres_mgsr_mf(kl.stvec, b, m, 1e-10, kl.no_precond, (0.0, 0.0))
    assert (f.restart_out, f.n_out) == (g.restart_out, g.n_out) or abs(((f.restart_out - 1) * m + f.n_out) - gi) <= 1
    gh = h.gmres_hh_dense(A, b, m, 1e-10)
    oh = ko.gmres_hh_dense(A, b, m, 1e-10)
    ghi = (gh.restart_out - 1) * m + gh.n_out
    print(f"gmres_hh_dense {ns}^2: gpu {ghi} oracle {oh.iterations}")
    assert gh.status == 0 and abs(ghi - oh.iterations) <= 1
    assert np.abs(gh.x - oh.x).max() < 1e-9
    assert gh.v_err[: gh.n_out].max() < 1e-26


def test_hilbert_drivers(kl, h, ko):
    # tests/test_hilbert.f90: b = matmul(H, 1), tol = 1e-15.  cond(H) ~ 1e16: only the residual estimate and
    # the size of the error are comparable, not iteration-by-iteration values.
    n, m = 40, 30
    H = h.generate_matrix(n)
    b = h.dense_matvec(H, np.ones(n))
    for name in ("gmres_mgsr_dense", "gmres_hh_dense"):
        g = getattr(h, name)(H, b, m, 1e-15)
        o = getattr(ko, name)(H, b, m, 1e-15)
        print(f"{name} hilbert {n}: gpu stages {g.restart_out} n_out {g.n_out} fe {g.final_err[g.n_out - 1]:.2e} "
              f"err {np.abs(g.x - 1).max():.2e} | oracle stages {o.restart_out} n_out {o.n_out} err {np.abs(o.x - 1).max():.2e}")
        assert g.status in (0, 1)
        assert np.linalg.norm(H @ g.x - b) / np.linalg.norm(b) < 1e-12
        assert np.abs(g.x - 1).max() < 10 * max(np.abs(o.x - 1).max(), 1e-4)


def test_hh_dense_with_m_up_to_n(kl, h, ko):
    """gmres_hh.f90:53 `if (j < n)`: m = n - 1 builds the last possible reflector, m = n takes the else-branch at
    j = n (H(j+1,j) = 0, no new reflector).  A generic non-symmetric matrix and tol = 0 force every step to run."""
    n = 12
    rng = np.random.default_rng(11)
    A = 4.0 * np.eye(n) + rng.standard_normal((n, n))
    b = rng.standard_normal(n)
    h.set_option(2, 2)          # two cycles
    try:
        for m in (n - 1, n):
            g = h.gmres_hh_dense(A, b, m, 0.0)
            o = ko.gmres_hh_dense(A, b, m, 0.0, max_stages=2)
            assert (g.n_out, g.restart_out) == (o.n_out, o.restart_out) == (m, 2), (m, g.n_out, g.restart_out, o.n_out)
            k = min(g.history.size, o.history.size, m)
            assert np.allclose(g.history[:k], o.history[:k], rtol=1e-9, atol=1e-15), m
            assert np.abs(g.x - o.x).max() < 1e-10 * np.abs(o.x).max()
            if m == n:
                assert np.linalg.norm(A @ g.x - b) < 1e-10 * np.linalg.norm(b)      # full Krylov space: exact solve
    finally:
        h.set_option(2, 1000)


def test_dense_errors(kl, h):
    A = np.eye(8)
    with pytest.raises(kl.KrylovError):
        h.gmres_hh_dense(A, np.ones(8), 9, 1e-10)      # m > n: v_j(j) out of bounds in the reference itself
