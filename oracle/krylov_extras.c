/*
 * krylov_extras.c -- CPU ORACLE, part 2 (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * CPU twins of the pieces BASELINE.json's north_star asks for that do NOT
 * exist in the reference (README.md:11 claims a Lanczos estimator, README.md:46
 * lists anisotropic diffusion as WIP; neither has code): the anisotropic
 * 5-point operator, a degree-k Chebyshev preconditioner and the Lanczos
 * spectral estimate.  PARITY UNPINNED: there is no reference behaviour to
 * match; these definitions are this repo's own (DESIGN.md "New components")
 * and the GPU kernels are checked against them.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef void (*ko_stencil_fn)(const double *x, double *y, int n);
typedef void (*ko_precond_fn)(ko_stencil_fn A_x, const double *r, double *z, double *aux,
                              const double *params, int n, int64_t len);

/* ---- anisotropic diffusion, constant coefficients, zero Dirichlet ------
 * y = 2(ex+ey) x - ex (x_{i-1} + x_{i+1}) - ey (x_{j+1} + x_{j-1})
 * evaluation order (shared with the CUDA kernel):
 *   sx = xl + xr ; sy = xd + xu ; t = fma(ex, sx, ey*sy) ; y = fma(cc, xc, -t)
 * missing neighbours count as 0. */
static double g_ex = 1.0, g_ey = 1.0;
void ko_set_aniso(double ex, double ey) { g_ex = ex; g_ey = ey; }
void ko_aniso(const double *x, double *y, int n)
{
    const int64_t N = n;
    const double ex = g_ex, ey = g_ey, cc = 2.0 * (ex + ey);
#pragma omp for collapse(2)
    for (int64_t j = 0; j < N; ++j)
        for (int64_t i = 0; i < N; ++i) {
            int64_t idx = i + j * N;
            double xl = i > 0 ? x[idx - 1] : 0.0, xr = i < N - 1 ? x[idx + 1] : 0.0;
            double xd = j < N - 1 ? x[idx + N] : 0.0, xu = j > 0 ? x[idx - N] : 0.0;
            double sx = xl + xr, sy = xd + xu;
            double t = fma(ex, sx, ey * sy);
            y[idx] = fma(cc, x[idx], -t);
        }
}
ko_stencil_fn ko_get_aniso(void) { return ko_aniso; }

/* ---- anisotropic diffusion with VARIABLE coefficients (README.md:46 WIP; definition ours) --------------
 * Self-adjoint finite-volume form of  -d/dx(kx du/dx) - d/dy(ky du/dy)  on the unit-spaced grid, zero Dirichlet:
 * cell coefficients kx(i,j), ky(i,j); a face carries the arithmetic mean of its two cells, a boundary face the
 * cell's own value.  Evaluation order (shared with the CUDA kernel k_aniso_var, kl_ops.cu):
 *   wl = 0.5*(kxc + kx(i-1,j))  [kxc at i = 0]   wr = 0.5*(kxc + kx(i+1,j))  [kxc at i = n-1]
 *   wu = 0.5*(kyc + ky(i,j-1))  [kyc at j = 0]   wd = 0.5*(kyc + ky(i,j+1))  [kyc at j = n-1]
 *   diag = ((wl + wr) + wd) + wu ; s = fma(wl, xl, wr*xr) ; t = fma(wd, xd, wu*xu) ; y = fma(diag, xc, -(s + t))
 * missing neighbours count as 0.  ("d" = idx+n, "u" = idx-n, the reference's naming in poisson.f90:42.) */
static const double *g_kx = NULL, *g_ky = NULL;
void ko_set_aniso_var(const double *kx, const double *ky) { g_kx = kx; g_ky = ky; }
void ko_aniso_var(const double *x, double *y, int n)
{
    const int64_t N = n;
    const double *kx = g_kx, *ky = g_ky;
#pragma omp for collapse(2)
    for (int64_t j = 0; j < N; ++j)
        for (int64_t i = 0; i < N; ++i) {
            int64_t idx = i + j * N;
            double kxc = kx[idx], kyc = ky[idx];
            double wl = i > 0 ? 0.5 * (kxc + kx[idx - 1]) : kxc, wr = i < N - 1 ? 0.5 * (kxc + kx[idx + 1]) : kxc;
            double wu = j > 0 ? 0.5 * (kyc + ky[idx - N]) : kyc, wd = j < N - 1 ? 0.5 * (kyc + ky[idx + N]) : kyc;
            double xl = i > 0 ? x[idx - 1] : 0.0, xr = i < N - 1 ? x[idx + 1] : 0.0;
            double xd = j < N - 1 ? x[idx + N] : 0.0, xu = j > 0 ? x[idx - N] : 0.0;
            double diag = ((wl + wr) + wd) + wu;
            double s = fma(wl, xl, wr * xr), t = fma(wd, xd, wu * xu);
            y[idx] = fma(diag, x[idx], -(s + t));
        }
}
ko_stencil_fn ko_get_aniso_var(void) { return ko_aniso_var; }

/* ---- degree-k Chebyshev preconditioner (Saad, Alg. 12.1) ---------------
 * params = (eig_a, eig_b) in either order (like cbpr2).
 *   theta = (b+a)/2, delta = |b-a|/2, sigma = theta/delta, rho_0 = 1/sigma
 *   d = r/theta ; z = d
 *   for i = 1..k: rho = 1/(2 sigma - rho_prev)
 *                 d = (rho*rho_prev) d + (2 rho/delta)(r - A z) ; z = z + d
 * evaluation order (shared with the CUDA kernel):
 *   d_new = fma(c1, d, c2*(r - Az)) ; z_new = z + d_new
 * needs TWO scratch vectors; aux holds A z, d is allocated here. */
static int g_cheb_degree = 1;
void ko_set_cheb_degree(int k) { g_cheb_degree = k; }
void ko_cheb(ko_stencil_fn A_x, const double *r, double *z, double *aux, const double *params,
             int n, int64_t len)
{
    static double *d = NULL;
    static int64_t dlen = 0;
#pragma omp single
    {
        if (dlen < len) { free(d); d = malloc(sizeof(double) * len); dlen = len; }
    }
    double ea = params[0], eb = params[1];
    double theta = (eb + ea) / 2.0, delta = fabs(eb - ea) / 2.0;
    double sigma = theta / delta, rho_prev = 1.0 / sigma;
#pragma omp for
    for (int64_t i = 0; i < len; ++i) { d[i] = r[i] / theta; z[i] = d[i]; }
    for (int k = 0; k < g_cheb_degree; ++k) {
        double rho = 1.0 / (2.0 * sigma - rho_prev);
        double c1 = rho * rho_prev, c2 = 2.0 * rho / delta;
        A_x(z, aux, n);
#pragma omp for
        for (int64_t i = 0; i < len; ++i) {
            d[i] = fma(c1, d[i], c2 * (r[i] - aux[i]));
            z[i] = z[i] + d[i];
        }
        rho_prev = rho;
    }
}
ko_precond_fn ko_get_cheb(void) { return ko_cheb; }

/* ---- Lanczos spectral estimate ------------------------------------------
 * k-step symmetric Lanczos on A, start vector b/||b|| with b = A*1 (the
 * drivers' right-hand side).  Returns the number of steps done; out[0] =
 * smallest, out[1] = largest Ritz value (bisection on the tridiagonal). */
static int sturm_count(const double *a, const double *b, int k, double x)
{
    /* number of eigenvalues of T_k smaller than x */
    int cnt = 0;
    double q = a[0] - x;
    if (q < 0) ++cnt;
    for (int i = 1; i < k; ++i) {
        double den = (q != 0.0) ? q : 1e-300;
        q = (a[i] - x) - b[i - 1] * b[i - 1] / den;
        if (q < 0) ++cnt;
    }
    return cnt;
}
void ko_tridiag_extremes(const double *a, const double *b, int k, double *lo_out, double *hi_out)
{
    double lo = a[0], hi = a[0];
    for (int i = 0; i < k; ++i) {
        double rad = (i > 0 ? fabs(b[i - 1]) : 0.0) + (i < k - 1 ? fabs(b[i]) : 0.0);
        if (a[i] - rad < lo) lo = a[i] - rad;
        if (a[i] + rad > hi) hi = a[i] + rad;
    }
    /* smallest: first x with count >= 1 ; largest: first x with count >= k */
    for (int which = 0; which < 2; ++which) {
        double l = lo, h = hi;
        int target = which == 0 ? 1 : k;
        for (int it = 0; it < 200; ++it) {
            double mid = 0.5 * (l + h);
            if (mid == l || mid == h) break;
            if (sturm_count(a, b, k, mid) >= target) h = mid; else l = mid;
        }
        if (which == 0) *lo_out = 0.5 * (l + h); else *hi_out = 0.5 * (l + h);
    }
}
int ko_lanczos(ko_stencil_fn A, int nsize, int steps, double *out, double *alphas, double *betas)
{
    int64_t n = (int64_t)nsize * nsize;
    double *v = malloc(sizeof(double) * n), *vp = calloc(n, sizeof(double));
    double *w = malloc(sizeof(double) * n);
    for (int64_t t = 0; t < n; ++t) vp[t] = 1.0;
    A(vp, v, nsize);
    double nb = 0.0;
    for (int64_t t = 0; t < n; ++t) nb = fma(v[t], v[t], nb);
    nb = sqrt(nb);
    for (int64_t t = 0; t < n; ++t) { v[t] = v[t] / nb; vp[t] = 0.0; }
    double beta_prev = 0.0;
    int k = 0;
    for (int i = 0; i < steps; ++i) {
        A(v, w, nsize);
        double al = 0.0;
        for (int64_t t = 0; t < n; ++t) al = fma(v[t], w[t], al);
        double bt = 0.0;
        for (int64_t t = 0; t < n; ++t) {
            w[t] = fma(-beta_prev, vp[t], fma(-al, v[t], w[t]));
            bt = fma(w[t], w[t], bt);
        }
        bt = sqrt(bt);
        alphas[i] = al; betas[i] = bt;
        k = i + 1;
        if (!(bt > 0.0)) break;
        for (int64_t t = 0; t < n; ++t) { vp[t] = v[t]; v[t] = w[t] / bt; }
        beta_prev = bt;
    }
    ko_tridiag_extremes(alphas, betas, k, &out[0], &out[1]);
    free(v); free(vp); free(w);
    return k;
}
