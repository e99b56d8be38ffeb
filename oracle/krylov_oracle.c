/*
 * krylov_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * A plain-C restatement of the reference's Fortran iterative-solver hot path
 * (AlexanderGSC/gmres, "Krylov Lab").  Every function cites the reference
 * file:line it follows.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library; the product
 * (libkrylov_b200.so) never links, loads or calls it.
 *
 * PINNING STATUS: the reference stores no golden vectors and no Fortran
 * compiler exists in the build container, so this oracle cannot be checked
 * against reference *output*.  It is pinned against what the reference's
 * drivers do assert implicitly (manufactured solution x == 1 with b = A*1,
 * ||b||_2 = sqrt(4*nsize+8)), against the README's quantitative claims
 * (Householder orthogonality ~1e-30 in calculate_verr's metric) and against
 * an independently written numpy restatement (tests/golden/).  For anything
 * beyond that: "parity unpinned".
 *
 * Arithmetic conventions (so that the restatement is well defined):
 *   - compiled with -ffp-contract=off; every place where gfortran -O3
 *     -march=native (no -ffast-math) would contract a*b+c into an FMA is
 *     written as an explicit fma() call.  Sums are never re-associated.
 *   - dot_product = sequential left-to-right FMA accumulation.
 *   - norm2 = libgfortran's scaled one-pass algorithm (norm2_r8).
 *   - OpenMP structure (parallel regions, orphaned work-sharing inside the
 *     operator and the preconditioner, single/master sections) mirrors the
 *     *_omp routines so that the same file doubles as the CPU baseline.  With
 *     OMP_NUM_THREADS=1 every reduction is a sequential sum (deterministic).
 *   - indices are 0-based here; "j" in comments is the reference's 1-based j.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef void (*ko_stencil_fn)(const double *x, double *y, int n);
typedef void (*ko_precond_fn)(ko_stencil_fn A_x, const double *r, double *z,
                              double *aux, const double *params, int n, int64_t len);

/* --------------------------------------------------------------------- */
/* helpers                                                               */
/* --------------------------------------------------------------------- */

/* libgfortran norm2_r8 (generated from m4/norm2.m4): scaled sum of squares. */
static double ko_norm2(const double *x, int64_t n)
{
    double scale = 1.0, result = 0.0;
    for (int64_t i = 0; i < n; ++i) {
        if (x[i] != 0.0) {
            double absx = fabs(x[i]);
            if (scale < absx) {
                double val = scale / absx;
                result = 1.0 + result * val * val;
                scale = absx;
            } else {
                double val = absx / scale;
                result += val * val;
            }
        }
    }
    return scale * sqrt(result);
}

/* Fortran dot_product on contiguous real(8): sequential FMA accumulation. */
static double ko_dot(const double *a, const double *b, int64_t n)
{
    double s = 0.0;
    for (int64_t i = 0; i < n; ++i) s = fma(a[i], b[i], s);
    return s;
}

/* nsize = int(sqrt(real(n))): SINGLE precision sqrt (gmres_mgsr.f90:298,
 * gmres_hh.f90:231, cg.f90:98, bicgstab.f90:109). */
static int ko_grid_side(int64_t n) { return (int)sqrtf((float)n); }

int ko_omp_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void ko_omp_set_threads(int t)
{
#ifdef _OPENMP
    omp_set_dynamic(0);
    omp_set_num_threads(t);
#else
    (void)t;
#endif
}

/* --------------------------------------------------------------------- */
/* src/problems/poisson.f90                                              */
/* --------------------------------------------------------------------- */

/* poisson.f90:33-77 stvec.  Column-major grid, idx = i + (j-1)*n, 1-based in
 * the reference; the neighbour-sum order of every piece is kept. */
void ko_stvec(const double *x, double *y, int n)
{
    const int64_t N = n;
    /* :38-45 interior, collapse(2) */
#pragma omp for collapse(2)
    for (int64_t j = 1; j < N - 1; ++j)
        for (int64_t i = 1; i < N - 1; ++i) {
            int64_t idx = i + j * N;
            y[idx] = 4.0 * x[idx] - (((x[idx - 1] + x[idx + 1]) + x[idx + N]) + x[idx - N]);
        }
    /* :46-49 col = 1 */
#pragma omp for nowait
    for (int64_t i = 1; i < N - 1; ++i)
        y[i] = 4.0 * x[i] - ((x[i - 1] + x[i + 1]) + x[i + N]);
    /* :50-55 col = n */
#pragma omp for
    for (int64_t i = 1; i < N - 1; ++i) {
        int64_t idx = N * N - N + i;
        y[idx] = 4.0 * x[idx] - ((x[idx - 1] + x[idx + 1]) + x[idx - N]);
    }
    /* :56-61 row = 1 */
#pragma omp for
    for (int64_t i = 1; i < N - 1; ++i) {
        int64_t idx = i * N;
        y[idx] = 4.0 * x[idx] - ((x[idx + 1] + x[idx + N]) + x[idx - N]);
    }
    /* :62-67 row = n */
#pragma omp for
    for (int64_t i = 1; i < N - 1; ++i) {
        int64_t idx = i * N + N - 1;
        y[idx] = 4.0 * x[idx] - ((x[idx - 1] + x[idx + N]) + x[idx - N]);
    }
    /* :69-76 corners */
#pragma omp single
    {
        int64_t idx;
        y[0] = 4.0 * x[0] - (x[1] + x[N]);
        y[N - 1] = 4.0 * x[N - 1] - (x[N - 2] + x[N - 1 + N]);
        idx = N * (N - 1);
        y[idx] = 4.0 * x[idx] - (x[idx + 1] + x[idx - N]);
        idx = N * (N - 1) + N - 1;
        y[idx] = 4.0 * x[idx] - (x[idx - 1] + x[idx - N]);
    }
}

/* poisson.f90:79-96 stv_poisson: branchy form, rounding order
 * ((((4x - x_{-1}) - x_{+1}) - x_{-n}) - x_{+n}).  "4.0"/"1.0" are default-real
 * literals promoted to double (exact). */
void ko_stv_poisson(const double *x, double *y, int n)
{
    const int64_t N = n;
#pragma omp for collapse(2)
    for (int64_t j = 0; j < N; ++j)
        for (int64_t i = 0; i < N; ++i) {
            int64_t idx = i + j * N;
            double v = 4.0 * x[idx];
            if (i > 0) v = v - x[idx - 1];
            if (i < N - 1) v = v - x[idx + 1];
            if (j > 0) v = v - x[idx - N];
            if (j < N - 1) v = v - x[idx + N];
            y[idx] = v;
        }
}

/* --------------------------------------------------------------------- */
/* src/preconds/chebyshev.f90                                            */
/* --------------------------------------------------------------------- */

/* chebyshev.f90:8-38 cbpr2.  len = size(r). */
void ko_cbpr2(ko_stencil_fn A_x, const double *r, double *z, double *aux,
              const double *params, int n, int64_t len)
{
    /* :19-26 (single + copyprivate): every thread computes the same values */
    double eigen_min = params[0], eigen_max = params[1];
    double c = (eigen_max - eigen_min) / 2.0;
    double d = (eigen_max + eigen_min) / 2.0;
    double alpha = 1.0 / d;
    double beta = (c * alpha / 2.0) * (c * alpha / 2.0);
    alpha = 1.0 / (d - beta);
    /* :27-31 */
#pragma omp for
    for (int64_t i = 0; i < len; ++i) z[i] = r[i] / d;
    /* :32 */
    A_x(z, aux, n);
    /* :33-37 */
#pragma omp for
    for (int64_t i = 0; i < len; ++i) z[i] = fma(alpha, r[i] - aux[i], z[i]);
}

/* identity "preconditioner" used to run the preconditioned entry points
 * unpreconditioned in tests (not in the reference). */
void ko_precond_identity(ko_stencil_fn A_x, const double *r, double *z, double *aux,
                         const double *params, int n, int64_t len)
{
    (void)A_x; (void)aux; (void)params; (void)n;
#pragma omp for
    for (int64_t i = 0; i < len; ++i) z[i] = r[i];
}

/* --------------------------------------------------------------------- */
/* Givens update shared by all GMRES flavours                            */
/* gmres_mgsr.f90:153-171 / :365-383, gmres_hh.f90:323-339 / :504-520      */
/* H is column-major (m+1) x m, ldh = m+1; j is 0-based column.           */
/* --------------------------------------------------------------------- */
static void ko_givens(double *H, int ldh, double *cs, double *sn, double *g, int j)
{
    double *Hj = H + (int64_t)j * ldh;
    for (int i = 0; i < j; ++i) {
        double tmp = Hj[i];
        Hj[i] = fma(cs[i], tmp, sn[i] * Hj[i + 1]);
        Hj[i + 1] = fma(-sn[i], tmp, cs[i] * Hj[i + 1]);
    }
    double ds = hypot(Hj[j + 1], Hj[j]);
    cs[j] = Hj[j] / ds;
    sn[j] = Hj[j + 1] / ds;
    Hj[j] = fma(cs[j], Hj[j], sn[j] * Hj[j + 1]);
    Hj[j + 1] = 0.0;
    double tmp = g[j];
    g[j] = fma(cs[j], tmp, sn[j] * g[j + 1]);
    g[j + 1] = fma(-sn[j], tmp, cs[j] * g[j + 1]);
}

/* back substitution, gmres_mgsr.f90:179-183 / :394-398, gmres_hh.f90:350-354.
 * dot_product(H(i,i+1:n_out), y(i+1:n_out)) sequential. */
static void ko_backsolve(const double *H, int ldh, const double *g, double *y, int m, int n_out)
{
    for (int i = 0; i < m; ++i) y[i] = 0.0;
    y[n_out - 1] = g[n_out - 1] / H[(int64_t)(n_out - 1) * ldh + (n_out - 1)];
    for (int i = n_out - 2; i >= 0; --i) {
        double s = 0.0;
        for (int k = i + 1; k < n_out; ++k) s = fma(H[(int64_t)k * ldh + i], y[k], s);
        y[i] = (g[i] - s) / H[(int64_t)i * ldh + i];
    }
}

/* orthogonality metric of the MGS solvers, gmres_mgsr.f90:192-198 / :414-420.
 * "2.0" is a default-real literal: 2.0*(dot**2) is exact scaling. */
static void ko_mgsr_verr(const double *V, int64_t n, int n_out, double *v_err)
{
    for (int j = 0; j < n_out; ++j) {            /* reference j = 1..n_out */
        const double *vj1 = V + (int64_t)(j + 1) * n;
        for (int i = 0; i <= j; ++i) {
            double d = ko_dot(V + (int64_t)i * n, vj1, n);
            v_err[j + 1] = v_err[j + 1] + 2.0 * (d * d);
        }
        double dd = ko_dot(vj1, vj1, n) - 1.0;
        v_err[j + 1] = v_err[j + 1] + dd * dd;
        v_err[j + 1] = sqrt(v_err[j] * v_err[j] + v_err[j + 1]);
    }
}

/* --------------------------------------------------------------------- */
/* src/gmres_mgsr.f90                                                    */
/* --------------------------------------------------------------------- */
#define KO_MAX_RESTARTS 1000  /* gmres_mgsr.f90:6 */
#define KO_HH_STAGES 1000     /* gmres_hh.f90:8  */

/* Optional outputs common to all solvers (not in the reference): history
 * receives one residual estimate per inner iteration, across restarts
 * (history_cap entries at most, *history_len = number produced). */

/* gmres_mgsr.f90:98-199 gmres_mgsr_mf (serial, early exit).
 * ortho: 0 = reference MGS x2 ; 1 = CGS2 (classical GS twice; NOT in the
 * reference -- provided so the GPU library's fast mode has a CPU twin). */
int ko_gmres_mgsr_mf(ko_stencil_fn Ax_vec, const double *b, int64_t n, double *x, int m,
                     double tol, double *final_err, double *v_err, int *n_out_p,
                     int *restart_out_p, ko_precond_fn M_inv, const double *params,
                     int max_restarts, int ortho, double *history, int history_cap,
                     int *history_len)
{
    int nsize = ko_grid_side(n);
    if (max_restarts <= 0) max_restarts = KO_MAX_RESTARTS;
    double *V = calloc((size_t)n * (m + 1), sizeof(double));
    double *H = calloc((size_t)(m + 1) * m, sizeof(double));
    double *y = calloc(m, sizeof(double)), *z = calloc(n, sizeof(double));
    double *aux = calloc(n, sizeof(double)), *w = calloc(n, sizeof(double));
    double *g = calloc(m + 1, sizeof(double));
    double *cs = calloc(m, sizeof(double)), *sn = calloc(m, sizeof(double));
    double *hh = calloc(m + 1, sizeof(double));
    if (!V || !H || !y || !z || !aux || !w || !g || !cs || !sn || !hh) return -1;
    const int ldh = m + 1;
    int n_out = 0, restart_out = max_restarts, hl = 0;
    double h_val = 0.0;
    memset(final_err, 0, sizeof(double) * m);
    memset(v_err, 0, sizeof(double) * (m + 1));
    memset(x, 0, sizeof(double) * n);
    double beta0 = ko_norm2(b, n); /* :125 */
    for (int st = 1; st <= max_restarts; ++st) {
        memset(g, 0, sizeof(double) * (m + 1));
        memset(H, 0, sizeof(double) * (size_t)(m + 1) * m);
        memset(V, 0, sizeof(double) * (size_t)n * (m + 1));
        Ax_vec(x, w, nsize);                                    /* :129 */
        for (int64_t i = 0; i < n; ++i) z[i] = b[i] - w[i];     /* :130 */
        M_inv(Ax_vec, z, w, aux, params, nsize, n);             /* :131 */
        double beta = ko_norm2(w, n);                           /* :132 */
        for (int64_t i = 0; i < n; ++i) V[i] = w[i] / beta;     /* :133 */
        g[0] = beta;
        for (int j = 0; j < m; ++j) {
            n_out = j + 1;
            double *Hj = H + (int64_t)j * ldh;
            Ax_vec(V + (int64_t)j * n, z, nsize);               /* :138 */
            M_inv(Ax_vec, z, w, aux, params, nsize, n);         /* :139 */
            for (int k = 0; k < 2; ++k) {                       /* :143-149 */
                if (ortho == 0) {
                    for (int i = 0; i <= j; ++i) {
                        const double *vi = V + (int64_t)i * n;
                        double h_tmp = ko_dot(w, vi, n);
                        Hj[i] = Hj[i] + h_tmp;
                        for (int64_t t = 0; t < n; ++t) w[t] = fma(-h_tmp, vi[t], w[t]);
                    }
                } else {
                    for (int i = 0; i <= j; ++i) hh[i] = ko_dot(w, V + (int64_t)i * n, n);
                    for (int i = 0; i <= j; ++i) {
                        const double *vi = V + (int64_t)i * n;
                        Hj[i] = Hj[i] + hh[i];
                        for (int64_t t = 0; t < n; ++t) w[t] = fma(-hh[i], vi[t], w[t]);
                    }
                }
            }
            h_val = ko_norm2(w, n);                             /* :150 */
            Hj[j + 1] = h_val;
            ko_givens(H, ldh, cs, sn, g, j);                    /* :153-168 */
            final_err[j] = fabs(g[j + 1]) / beta0;              /* :171 */
            if (history && hl < history_cap) history[hl] = final_err[j];
            ++hl;
            if (h_val < tol || final_err[j] < tol) {            /* :172-175 */
                n_out = j + 1;
                break;
            }
            double *vj1 = V + (int64_t)(j + 1) * n;             /* :176 */
            for (int64_t t = 0; t < n; ++t) vj1[t] = w[t] / h_val;
        }
        ko_backsolve(H, ldh, g, y, m, n_out);                   /* :179-183 */
        /* :185 x = x + matmul(V(:,1:n_out),y): gfortran inlines matmul as a
         * column-sweep (axpy per column, FMA), then adds to x. */
        for (int64_t t = 0; t < n; ++t) w[t] = 0.0;
        for (int k = 0; k < n_out; ++k) {
            const double *vk = V + (int64_t)k * n;
            for (int64_t t = 0; t < n; ++t) w[t] = fma(vk[t], y[k], w[t]);
        }
        for (int64_t t = 0; t < n; ++t) x[t] = x[t] + w[t];
        if (h_val < tol || final_err[n_out - 1] < tol) {        /* :187-190 */
            restart_out = st;
            break;
        }
    }
    ko_mgsr_verr(V, n, n_out, v_err);                           /* :192-198 */
    *n_out_p = n_out;
    *restart_out_p = restart_out;
    if (history_len) *history_len = hl;
    free(V); free(H); free(y); free(z); free(aux); free(w); free(g); free(cs); free(sn); free(hh);
    return 0;
}

/* gmres_mgsr.f90:277-421 gmres_mgsr_omp.  ortho as above. */
int ko_gmres_mgsr_omp(ko_stencil_fn Ax_vec, const double *b, int64_t n, double *x, int m,
                      double tol, double *final_err, double *v_err, int *n_out_p,
                      int *restart_out_p, ko_precond_fn M_inv, const double *params,
                      int max_restarts, int ortho, int skip_verr, double *history,
                      int history_cap, int *history_len)
{
    int nsize = ko_grid_side(n);
    if (max_restarts <= 0) max_restarts = KO_MAX_RESTARTS;
    double *V = malloc(sizeof(double) * (size_t)n * (m + 1));
    double *H = calloc((size_t)(m + 1) * m, sizeof(double));
    double *y = calloc(m, sizeof(double)), *z = calloc(n, sizeof(double));
    double *aux = calloc(n, sizeof(double)), *w = calloc(n, sizeof(double));
    double *g = calloc(m + 1, sizeof(double));
    double *cs = calloc(m, sizeof(double)), *sn = calloc(m, sizeof(double));
    double *hh = calloc(m + 1, sizeof(double));
    if (!V || !H || !y || !z || !aux || !w || !g || !cs || !sn || !hh) return -1;
    const int ldh = m + 1;
    /* shared state of the parallel team */
    int converged = 0, n_out = 0, restart_out = max_restarts, hl = 0;
    double h_val = 0.0, h_tmp = 0.0, beta = 0.0;
    memset(final_err, 0, sizeof(double) * m);
    memset(v_err, 0, sizeof(double) * (m + 1));
    memset(x, 0, sizeof(double) * n);
    double beta0 = ko_norm2(b, n); /* :307 */
    for (int st = 1; st <= max_restarts; ++st) {
#pragma omp parallel
        {
            /* :311-313 workshare g=0;H=0;V=0 */
#pragma omp single
            {
                memset(g, 0, sizeof(double) * (m + 1));
                memset(H, 0, sizeof(double) * (size_t)(m + 1) * m);
            }
#pragma omp for
            for (int64_t t = 0; t < (int64_t)n * (m + 1); ++t) V[t] = 0.0;
            Ax_vec(x, w, nsize);                                        /* :314 */
#pragma omp for
            for (int64_t t = 0; t < n; ++t) z[t] = b[t] - w[t];         /* :315-319 */
            M_inv(Ax_vec, z, w, aux, params, nsize, n);                 /* :320 */
#pragma omp single
            {
                beta = ko_norm2(w, n);                                  /* :322 */
                g[0] = beta;
            }
#pragma omp for
            for (int64_t t = 0; t < n; ++t) V[t] = w[t] / beta;         /* :325-329 */
        }
#pragma omp parallel
        {
            for (int j = 0; j < m; ++j) {
                if (converged) continue;                                /* :335 */
                double *Hj = H + (int64_t)j * ldh;
                Ax_vec(V + (int64_t)j * n, z, nsize);                   /* :336 */
                M_inv(Ax_vec, z, w, aux, params, nsize, n);             /* :337 */
                for (int k = 0; k < 2; ++k) {                           /* :341 */
                    if (ortho == 0) {
                        for (int i = 0; i <= j; ++i) {
                            const double *vi = V + (int64_t)i * n;
#pragma omp single
                            h_tmp = 0.0;
#pragma omp for reduction(+ : h_tmp)
                            for (int64_t t = 0; t < n; ++t) h_tmp = fma(w[t], vi[t], h_tmp);
#pragma omp master
                            Hj[i] = Hj[i] + h_tmp;
                            /* (the reference has no barrier after master; the
                             * following omp-do only reads h_tmp) */
#pragma omp for
                            for (int64_t t = 0; t < n; ++t) w[t] = fma(-h_tmp, vi[t], w[t]);
                        }
                    } else {
                        /* CGS pass: all projections from the same w, then one update */
                        for (int i = 0; i <= j; ++i) {
                            const double *vi = V + (int64_t)i * n;
#pragma omp single
                            h_tmp = 0.0;
#pragma omp for reduction(+ : h_tmp)
                            for (int64_t t = 0; t < n; ++t) h_tmp = fma(w[t], vi[t], h_tmp);
#pragma omp single
                            {
                                hh[i] = h_tmp;
                                Hj[i] = Hj[i] + h_tmp;
                            }
                        }
#pragma omp for
                        for (int64_t t = 0; t < n; ++t) {
                            double wt = w[t];
                            for (int i = 0; i <= j; ++i) wt = fma(-hh[i], V[(int64_t)i * n + t], wt);
                            w[t] = wt;
                        }
                    }
                }
#pragma omp single
                {
                    h_val = ko_norm2(w, n);                             /* :362 */
                    Hj[j + 1] = h_val;
                    ko_givens(H, ldh, cs, sn, g, j);                    /* :365-380 */
                    final_err[j] = fabs(g[j + 1]) / beta0;              /* :383 */
                    if (history && hl < history_cap) history[hl] = final_err[j];
                    ++hl;
                    double *vj1 = V + (int64_t)(j + 1) * n;             /* :384 */
                    for (int64_t t = 0; t < n; ++t) vj1[t] = w[t] / h_val;
                    if (final_err[j] < tol) {                           /* :385-388 */
                        restart_out = st;
                        converged = 1;
                    }
                    n_out = j + 1;                                      /* :389 */
                }
            }
        }
        ko_backsolve(H, ldh, g, y, m, n_out);                           /* :394-398 */
#pragma omp parallel for
        for (int64_t t = 0; t < n; ++t) {                               /* :400-406 */
            double s = 0.0;
            for (int k = 0; k < n_out; ++k) s = fma(V[(int64_t)k * n + t], y[k], s);
            x[t] = x[t] + s;
        }
        if (h_val < tol || final_err[n_out - 1] < tol) {                /* :409-412 */
            restart_out = st;
            break;
        }
    }
    if (!skip_verr) ko_mgsr_verr(V, n, n_out, v_err);                   /* :414-420 */
    *n_out_p = n_out;
    *restart_out_p = restart_out;
    if (history_len) *history_len = hl;
    free(V); free(H); free(y); free(z); free(aux); free(w); free(g); free(cs); free(sn); free(hh);
    return 0;
}

/* --------------------------------------------------------------------- */
/* src/gmres_hh.f90                                                      */
/* --------------------------------------------------------------------- */

/* gmres_hh.f90:568-593 calculate_verr.  v_err is inout (accumulates). x is
 * overwritten with V*y (callers pass a scratch vector). */
void ko_calculate_verr(const double *P, int64_t n, double *x, const double *y, double *v_err,
                       int n_iter)
{
    double *V = calloc((size_t)n * n_iter, sizeof(double));
    for (int i = 0; i < n_iter; ++i) V[(int64_t)i * n + i] = 1.0;       /* :578-580 */
    for (int i = 0; i < n_iter; ++i) {                                  /* :581-585 */
        double *vi = V + (int64_t)i * n;
        for (int j = i; j >= 0; --j) {
            const double *pj = P + (int64_t)j * n;
            double d = ko_dot(vi, pj, n);
            /* V(:,i) - 2.0d0*P(:,j)*dot : (2P) exact, then fused multiply-sub */
            for (int64_t t = 0; t < n; ++t) vi[t] = fma(-(2.0 * pj[t]), d, vi[t]);
        }
    }
    for (int i = 1; i < n_iter; ++i)                                    /* :587-591 */
        for (int j = 0; j < i; ++j) {
            double d = ko_dot(V + (int64_t)i * n, V + (int64_t)j * n, n);
            v_err[i] = v_err[i] + 2.0 * (d * d);
        }
    /* :592 x = matmul(V, y(1:n_iter)) */
    for (int64_t t = 0; t < n; ++t) x[t] = 0.0;
    for (int k = 0; k < n_iter; ++k) {
        const double *vk = V + (int64_t)k * n;
        for (int64_t t = 0; t < n; ++t) x[t] = fma(vk[t], y[k], x[t]);
    }
    free(V);
}

/* ||I - V^T V||_F of the Arnoldi basis rebuilt from the reflectors (not in
 * the reference; used to compare orthogonality levels). */
double ko_hh_orth_frobenius(const double *P, int64_t n, int n_iter)
{
    double *V = calloc((size_t)n * n_iter, sizeof(double));
    for (int i = 0; i < n_iter; ++i) V[(int64_t)i * n + i] = 1.0;
    for (int i = 0; i < n_iter; ++i) {
        double *vi = V + (int64_t)i * n;
        for (int j = i; j >= 0; --j) {
            const double *pj = P + (int64_t)j * n;
            double d = ko_dot(vi, pj, n);
            for (int64_t t = 0; t < n; ++t) vi[t] = fma(-(2.0 * pj[t]), d, vi[t]);
        }
    }
    double s = 0.0;
    for (int i = 0; i < n_iter; ++i)
        for (int j = 0; j < n_iter; ++j) {
            double d = ko_dot(V + (int64_t)i * n, V + (int64_t)j * n, n) - (i == j ? 1.0 : 0.0);
            s += d * d;
        }
    free(V);
    return sqrt(s);
}

/* shared serial block of the Householder solvers,
 * gmres_hh.f90:305-339 / :486-520 (0-based j). */
static void ko_hh_single_block(double *H, int ldh, double *P, int64_t n, double *w, double *cs,
                               double *sn, double *g, int j, double *h_val)
{
    double *Hj = H + (int64_t)j * ldh;
    for (int i = 0; i <= j; ++i) Hj[i] = w[i];                          /* :306 */
    if (j + 1 < n) {                                                    /* :307 */
        double tmp = ko_norm2(w + j + 1, n - (j + 1));                  /* :308 */
        Hj[j + 1] = (w[j + 1] > 0.0) ? -tmp : tmp;                      /* :309-313 */
        *h_val = fabs(Hj[j + 1]);
        for (int i = 0; i <= j; ++i) w[i] = 0.0;                        /* :315 */
        w[j + 1] = w[j + 1] - Hj[j + 1];                                /* :316 */
        double nw = ko_norm2(w, n);                                     /* :317 */
        double *pj1 = P + (int64_t)(j + 1) * n;
        for (int64_t t = 0; t < n; ++t) {
            w[t] = w[t] / nw;
            pj1[t] = w[t];                                              /* :318 */
        }
    } else {
        Hj[j + 1] = 0.0;
    }
    ko_givens(H, ldh, cs, sn, g, j);                                    /* :323-337 */
}

/* one reflector application v -= 2 P_i (P_i . v) with the reference's omp
 * structure (gmres_hh.f90:269-283).  Must be called by the whole team.
 * dotsum points at a team-shared scalar. */
static void ko_hh_reflect_team(double *v, const double *p, int64_t n, double *dotsum)
{
#pragma omp single
    *dotsum = 0.0;
#pragma omp for reduction(+ : dotsum[:1])
    for (int64_t t = 0; t < n; ++t) dotsum[0] = fma(v[t], p[t], dotsum[0]);
    double d = *dotsum;
#pragma omp for
    for (int64_t t = 0; t < n; ++t) v[t] = fma(-(2.0 * p[t]), d, v[t]);
}

/* gmres_hh.f90:211-385 gmres_hh_omp (M_inv == NULL) and
 * gmres_hh.f90:388-566 gmres_hh_prec_omp (M_inv != NULL). */
int ko_gmres_hh(ko_stencil_fn Ax_vec, const double *b, int64_t n, double *x, int m, double tol,
                double *final_err, double *v_err, int *n_out_p, int *stages_out_p,
                ko_precond_fn M_inv, const double *params, int max_stages, int skip_verr,
                double *orth_frob, double *history, int history_cap, int *history_len)
{
    int nsize = ko_grid_side(n);
    if (max_stages <= 0) max_stages = KO_HH_STAGES;
    double *P = malloc(sizeof(double) * (size_t)n * (m + 1));
    double *H = calloc((size_t)(m + 1) * m, sizeof(double));
    double *y = calloc(m, sizeof(double)), *v_j = calloc(n, sizeof(double));
    double *w = calloc(n, sizeof(double)), *g = calloc(m + 1, sizeof(double));
    double *cs = calloc(m, sizeof(double)), *sn = calloc(m, sizeof(double));
    double *z = calloc(n, sizeof(double)), *aux = calloc(n, sizeof(double));
    if (!P || !H || !y || !v_j || !w || !g || !cs || !sn || !z || !aux) return -1;
    const int ldh = m + 1;
    int converged = 0, n_out = 0, stages_out = 0, hl = 0;
    double h_val = 0.0, dotsum = 0.0;
    memset(x, 0, sizeof(double) * n);
    memset(final_err, 0, sizeof(double) * m);
    memset(v_err, 0, sizeof(double) * (m + 1));
    double beta0 = ko_norm2(b, n);                                      /* :237 / :419 */
    for (int k = 1; k <= max_stages; ++k) {
#pragma omp parallel
        {
#pragma omp single
            {
                memset(g, 0, sizeof(double) * (m + 1));
                memset(H, 0, sizeof(double) * (size_t)(m + 1) * m);
            }
#pragma omp for
            for (int64_t t = 0; t < (int64_t)n * (m + 1); ++t) P[t] = 0.0;   /* :241 */
            Ax_vec(x, w, nsize);                                        /* :243 */
            if (!M_inv) {
#pragma omp for
                for (int64_t t = 0; t < n; ++t) w[t] = b[t] - w[t];     /* :244-248 */
            } else {
#pragma omp for
                for (int64_t t = 0; t < n; ++t) z[t] = b[t] - w[t];     /* :426-430 */
                M_inv(Ax_vec, z, w, aux, params, nsize, n);             /* :431 */
            }
#pragma omp single
            {
                double beta = ko_norm2(w, n);                           /* :250 / :433 */
                g[0] = -copysign(beta, w[0]);                           /* :251 */
                w[0] = copysign(beta, w[0]) + w[0];                     /* :252 */
                double nw = ko_norm2(w, n);
                for (int64_t t = 0; t < n; ++t) P[t] = w[t] / nw;       /* :253 */
            }
            for (int j = 0; j < m; ++j) {
                if (M_inv && converged) continue;                       /* :439 (prec only) */
#pragma omp for
                for (int64_t t = 0; t < n; ++t) v_j[t] = 0.0;           /* :257-261 */
#pragma omp single
                {
                    n_out = j + 1;
                    v_j[j] = 1.0;
                }
                for (int i = j; i >= 0; --i)                            /* :269-283 */
                    ko_hh_reflect_team(v_j, P + (int64_t)i * n, n, &dotsum);
                if (!M_inv) {
                    Ax_vec(v_j, w, nsize);                              /* :285 */
                } else {
                    Ax_vec(v_j, z, nsize);                              /* :469 */
                    M_inv(Ax_vec, z, w, aux, params, nsize, n);         /* :470 */
                }
                for (int i = 0; i <= j; ++i)                            /* :290-304 */
                    ko_hh_reflect_team(w, P + (int64_t)i * n, n, &dotsum);
#pragma omp single
                {
                    ko_hh_single_block(H, ldh, P, n, w, cs, sn, g, j, &h_val);
                    final_err[j] = fabs(g[j + 1]) / beta0;              /* :339 / :520 */
                    if (history && hl < history_cap) history[hl] = final_err[j];
                    ++hl;
                    if (M_inv && final_err[j] < tol) {                  /* :521-525 */
                        n_out = j + 1;
                        stages_out = k;
                        converged = 1;
                    }
                }
            }
        }
        ko_backsolve(H, ldh, g, y, m, n_out);                           /* :350-354 */
        for (int64_t t = 0; t < n; ++t) w[t] = 0.0;                     /* :356-357 */
        for (int i = 0; i < n_out; ++i) w[i] = y[i];
        for (int i = n_out - 1; i >= 0; --i) {                          /* :361-373 */
            const double *p = P + (int64_t)i * n;
            double s = 0.0;
#pragma omp parallel for reduction(+ : s)
            for (int64_t t = 0; t < n; ++t) s = fma(w[t], p[t], s);
#pragma omp parallel for
            for (int64_t t = 0; t < n; ++t) w[t] = fma(-(2.0 * p[t]), s, w[t]);
        }
#pragma omp parallel for
        for (int64_t t = 0; t < n; ++t) x[t] = x[t] + w[t];             /* :374-378 */
        stages_out = k;                                                 /* :381 */
        if (final_err[n_out - 1] < tol) break;                          /* :382 */
    }
    if (orth_frob) *orth_frob = ko_hh_orth_frobenius(P, n, n_out);
    if (!skip_verr) ko_calculate_verr(P, n, w, y, v_err, n_out);        /* :384 / :565 */
    *n_out_p = n_out;
    *stages_out_p = stages_out;
    if (history_len) *history_len = hl;
    free(P); free(H); free(y); free(v_j); free(w); free(g); free(cs); free(sn); free(z); free(aux);
    return 0;
}

/* --------------------------------------------------------------------- */
/* dense-operator variants: gmres_mgsr_dense (src/gmres_mgsr.f90:11-95),  */
/* gmres_hh_dense (src/gmres_hh.f90:10-112), hilbert::generate_matrix      */
/* (src/problems/hilbert.f90:6-18)                                         */
/* --------------------------------------------------------------------- */
/* A is column-major n x n like the Fortran array.  matmul(A, v): libgfortran's
 * matmul_r8 blocks and unrolls this product in a way that cannot be restated
 * without the library; the oracle (and the CUDA kernel) define it as the
 * column sweep  y = 0 ; do j: y(:) = y(:) + A(:,j)*v(j)  with FMA, i.e. every
 * y(i) is a sequential sum over j.  PARITY UNPINNED beyond that. */
static const double *g_dense_A = NULL;
static int64_t g_dense_n = 0;
void ko_set_dense(const double *A, int64_t n) { g_dense_A = A; g_dense_n = n; }
void ko_dense_matvec(const double *x, double *y, int nsize_unused)
{
    (void)nsize_unused;
    const int64_t n = g_dense_n;
    const double *A = g_dense_A;
    for (int64_t i = 0; i < n; ++i) y[i] = 0.0;
    for (int64_t j = 0; j < n; ++j) {
        const double xj = x[j];
        const double *col = A + j * n;
        for (int64_t i = 0; i < n; ++i) y[i] = fma(col[i], xj, y[i]);
    }
}
ko_stencil_fn ko_get_dense(void) { return ko_dense_matvec; }

/* hilbert.f90:13-17  H(i,j) = 1 / real(i+j-1): real() is default (single)
 * precision, so the quotient is a float that is then widened to double. */
void ko_generate_matrix(double *H, int n)
{
    for (int j = 1; j <= n; ++j)
        for (int i = 1; i <= n; ++i) H[(int64_t)(i - 1) + (int64_t)(j - 1) * n] = (double)(1.0f / (float)(i + j - 1));
}

/* gmres_hh.f90:10-112 gmres_hh_dense: serial Householder GMRES on the dense
 * operator set by ko_set_dense; in-cycle exit on h_val < tol or final_err < tol. */
int ko_gmres_hh_dense(const double *b, int64_t n, double *x, int m, double tol, double *final_err,
                      double *v_err, int *n_out_p, int *stages_out_p, int max_stages,
                      double *history, int history_cap, int *history_len)
{
    if (max_stages <= 0) max_stages = KO_HH_STAGES;
    double *P = malloc(sizeof(double) * (size_t)n * (m + 1));
    double *H = calloc((size_t)(m + 1) * m, sizeof(double));
    double *y = calloc(m, sizeof(double)), *v_j = calloc(n, sizeof(double));
    double *w = calloc(n, sizeof(double)), *g = calloc(m + 1, sizeof(double));
    double *cs = calloc(m, sizeof(double)), *sn = calloc(m, sizeof(double));
    if (!P || !H || !y || !v_j || !w || !g || !cs || !sn) return -1;
    const int ldh = m + 1;
    int n_out = 0, stages_out = max_stages, hl = 0;
    double h_val = 0.0;
    memset(x, 0, sizeof(double) * n);
    memset(final_err, 0, sizeof(double) * m);
    memset(v_err, 0, sizeof(double) * (m + 1));
    double beta0 = ko_norm2(b, n);                                      /* :34 */
    for (int k = 1; k <= max_stages; ++k) {
        memset(g, 0, sizeof(double) * (m + 1));                         /* :36 */
        memset(H, 0, sizeof(double) * (size_t)(m + 1) * m);
        memset(P, 0, sizeof(double) * (size_t)n * (m + 1));
        ko_dense_matvec(x, w, 0);                                       /* :37 */
        for (int64_t t = 0; t < n; ++t) w[t] = b[t] - w[t];
        double beta = ko_norm2(w, n);                                   /* :38 */
        g[0] = -copysign(beta, w[0]);                                   /* :39 */
        w[0] = copysign(beta, w[0]) + w[0];                             /* :40 */
        double nw = ko_norm2(w, n);
        for (int64_t t = 0; t < n; ++t) P[t] = w[t] / nw;               /* :41 */
        int stop = 0;
        for (int j = 0; j < m; ++j) {
            n_out = j + 1;
            for (int64_t t = 0; t < n; ++t) v_j[t] = 0.0;               /* :44 */
            v_j[j] = 1.0;
            for (int i = j; i >= 0; --i) {                              /* :45-47 */
                const double *p = P + (int64_t)i * n;
                double d = ko_dot(v_j, p, n);
                for (int64_t t = 0; t < n; ++t) v_j[t] = fma(-(2.0 * p[t]), d, v_j[t]);
            }
            ko_dense_matvec(v_j, w, 0);                                 /* :48 */
            for (int i = 0; i <= j; ++i) {                              /* :49-51 */
                const double *p = P + (int64_t)i * n;
                double d = ko_dot(w, p, n);
                for (int64_t t = 0; t < n; ++t) w[t] = fma(-(2.0 * p[t]), d, w[t]);
            }
            ko_hh_single_block(H, ldh, P, n, w, cs, sn, g, j, &h_val);  /* :52-85 */
            final_err[j] = fabs(g[j + 1]) / beta0;                      /* :87 */
            if (history && hl < history_cap) history[hl] = final_err[j];
            ++hl;
            if (h_val < tol || final_err[j] < tol) {                    /* :88-92 */
                n_out = j + 1;
                stages_out = k;
                stop = 1;
                break;
            }
        }
        ko_backsolve(H, ldh, g, y, m, n_out);                           /* :95-99 */
        for (int64_t t = 0; t < n; ++t) w[t] = 0.0;                     /* :101-102 */
        for (int i = 0; i < n_out; ++i) w[i] = y[i];
        for (int i = n_out - 1; i >= 0; --i) {                          /* :103-105 */
            const double *p = P + (int64_t)i * n;
            double d = ko_dot(p, w, n);
            for (int64_t t = 0; t < n; ++t) w[t] = fma(-(2.0 * p[t]), d, w[t]);
        }
        for (int64_t t = 0; t < n; ++t) x[t] = x[t] + w[t];             /* :106 */
        if (stop || h_val < tol || final_err[n_out - 1] < tol) {        /* :108-111 */
            stages_out = k;
            break;
        }
    }
    ko_calculate_verr(P, n, w, y, v_err, n_out);                        /* :113 */
    *n_out_p = n_out;
    *stages_out_p = stages_out;
    if (history_len) *history_len = hl;
    free(P); free(H); free(y); free(v_j); free(w); free(g); free(cs); free(sn);
    return 0;
}

/* --------------------------------------------------------------------- */
/* src/cg.f90                                                            */
/* --------------------------------------------------------------------- */

/* cg.f90:11-42 cg (M_inv == NULL) and cg.f90:44-81 pcg (M_inv != NULL); serial.
 * *iter: max on entry, count on exit (unchanged if not converged). */
int ko_cg_serial(ko_stencil_fn Ax_op, const double *b, int64_t n, double *x, double tol,
                 int *iter, double *res_p, ko_precond_fn M_inv, const double *params,
                 double *history, int history_cap, int *history_len)
{
    int grid = ko_grid_side(n), hl = 0;
    double *ax = calloc(n, sizeof(double)), *p = calloc(n, sizeof(double));
    double *r = calloc(n, sizeof(double)), *z = NULL, *aux = NULL;
    double res = 0.0;
    if (M_inv) { z = calloc(n, sizeof(double)); aux = calloc(n, sizeof(double)); }
    for (int64_t t = 0; t < n; ++t) { x[t] = 0.0; r[t] = b[t]; }
    if (M_inv) {
        M_inv(Ax_op, r, z, aux, params, grid, n);                       /* :64 */
        for (int64_t t = 0; t < n; ++t) p[t] = z[t];
    } else {
        for (int64_t t = 0; t < n; ++t) p[t] = r[t];
    }
    const double *zz = M_inv ? z : r;
    int maxit = *iter;
    for (int i = 1; i <= maxit; ++i) {
        Ax_op(p, ax, grid);                                             /* :29 / :67 */
        double rr = ko_dot(r, zz, n);                                   /* :30 / :68 */
        double alpha = rr / ko_dot(ax, p, n);                           /* :31 */
        for (int64_t t = 0; t < n; ++t) x[t] = fma(alpha, p[t], x[t]);  /* :32 */
        for (int64_t t = 0; t < n; ++t) r[t] = fma(-alpha, ax[t], r[t]);/* :33 */
        res = ko_norm2(r, n);                                           /* :34 */
        if (history && hl < history_cap) history[hl] = res;
        ++hl;
        if (M_inv) M_inv(Ax_op, r, z, aux, params, grid, n);            /* :73 */
        double beta = ko_dot(r, zz, n) / rr;                            /* :35 / :74 */
        for (int64_t t = 0; t < n; ++t) p[t] = fma(beta, p[t], zz[t]);  /* :36 / :75 */
        if (res < tol) { *iter = i; break; }                            /* :37-40 */
    }
    *res_p = res;
    if (history_len) *history_len = hl;
    free(ax); free(p); free(r); free(z); free(aux);
    return 0;
}

/* cg.f90:83-152 cg_omp (M_inv == NULL) and cg.f90:154-234 pcg_omp. */
int ko_cg_omp(ko_stencil_fn Ax_op, const double *b, int64_t n, double *x, double tol, int *iter,
              double *res_p, ko_precond_fn M_inv, const double *params, double *history,
              int history_cap, int *history_len)
{
    int grid = ko_grid_side(n), hl = 0;
    double *ax = malloc(sizeof(double) * n), *p = malloc(sizeof(double) * n);
    double *r = malloc(sizeof(double) * n), *z = NULL, *aux = NULL;
    if (M_inv) { z = malloc(sizeof(double) * n); aux = malloc(sizeof(double) * n); }
    int converged = 0, maxit = *iter;
    double alpha = 0.0, beta = 0.0, rr = 0.0, res = 0.0;
#pragma omp parallel
    {
        if (!M_inv) {
#pragma omp for
            for (int64_t t = 0; t < n; ++t) { x[t] = 0.0; r[t] = b[t]; p[t] = b[t]; } /* :102-108 */
        } else {
#pragma omp for
            for (int64_t t = 0; t < n; ++t) { x[t] = 0.0; r[t] = b[t]; }              /* :176-181 */
            M_inv(Ax_op, r, z, aux, params, grid, n);                                 /* :182 */
#pragma omp for
            for (int64_t t = 0; t < n; ++t) p[t] = z[t];                              /* :183-187 */
        }
        const double *zz = M_inv ? z : r;
        for (int i = 1; i <= maxit; ++i) {
            if (converged) continue;                                    /* :110 / :189 */
            Ax_op(p, ax, grid);                                         /* :111 / :190 */
#pragma omp single
            { rr = 0.0; alpha = 0.0; res = 0.0; beta = 0.0; }           /* :112-117 */
#pragma omp for reduction(+ : rr, alpha)
            for (int64_t t = 0; t < n; ++t) {                           /* :118-123 / :197-202 */
                rr = fma(r[t], zz[t], rr);
                alpha = fma(ax[t], p[t], alpha);
            }
#pragma omp single
            alpha = rr / alpha;                                         /* :124-126 */
            if (!M_inv) {
#pragma omp for reduction(+ : res, beta)
                for (int64_t t = 0; t < n; ++t) {                       /* :127-134 */
                    x[t] = fma(alpha, p[t], x[t]);
                    r[t] = fma(-alpha, ax[t], r[t]);
                    res = fma(r[t], r[t], res);
                    beta = fma(r[t], r[t], beta);
                }
#pragma omp single
                { res = sqrt(res); beta = beta / rr; }                  /* :135-138 */
#pragma omp for
                for (int64_t t = 0; t < n; ++t) p[t] = fma(beta, p[t], r[t]); /* :139-143 */
#pragma omp single
                {
                    if (history && hl < history_cap) history[hl] = res;
                    ++hl;
                    if (res < tol) { converged = 1; *iter = i; }        /* :144-149 */
                }
            } else {
#pragma omp for reduction(+ : res)
                for (int64_t t = 0; t < n; ++t) {                       /* :206-212 */
                    x[t] = fma(alpha, p[t], x[t]);
                    r[t] = fma(-alpha, ax[t], r[t]);
                    res = fma(r[t], r[t], res);
                }
                M_inv(Ax_op, r, z, aux, params, grid, n);               /* :213 */
#pragma omp for reduction(+ : beta)
                for (int64_t t = 0; t < n; ++t) beta = fma(r[t], z[t], beta); /* :214-218 */
#pragma omp single
                {
                    res = sqrt(res);                                    /* :219-226 */
                    beta = beta / rr;
                    if (history && hl < history_cap) history[hl] = res;
                    ++hl;
                    if (res < tol) { converged = 1; *iter = i; }
                }
#pragma omp for
                for (int64_t t = 0; t < n; ++t) p[t] = fma(beta, p[t], z[t]); /* :227-231 */
            }
        }
    }
    *res_p = res;
    if (history_len) *history_len = hl;
    free(ax); free(p); free(r); free(z); free(aux);
    return 0;
}

/* --------------------------------------------------------------------- */
/* src/bicgstab.f90                                                      */
/* --------------------------------------------------------------------- */

/* bicgstab.f90:12-47 bicgstab (M_inv == NULL), :49-89 pbicgstab; serial. */
int ko_bicgstab_serial(ko_stencil_fn ax_op, const double *b, int64_t n, double *x, double tol,
                       int *iter, double *res_p, ko_precond_fn m_inv, const double *params,
                       double *history, int history_cap, int *history_len)
{
    int grid = ko_grid_side(n), hl = 0;
    double *r = calloc(n, sizeof(double)), *r0 = calloc(n, sizeof(double));
    double *ap = calloc(n, sizeof(double)), *s = calloc(n, sizeof(double));
    double *as = calloc(n, sizeof(double)), *p = calloc(n, sizeof(double));
    double *z1 = NULL, *z2 = NULL, *aux = NULL;
    if (m_inv) { z1 = calloc(n, sizeof(double)); z2 = calloc(n, sizeof(double)); aux = calloc(n, sizeof(double)); }
    double res = 0.0;
    for (int64_t t = 0; t < n; ++t) { x[t] = 0.0; r[t] = b[t]; r0[t] = b[t]; p[t] = b[t]; }
    int maxit = *iter;
    for (int i = 1; i <= maxit; ++i) {
        const double *pp = p, *ss = s;
        if (m_inv) { m_inv(ax_op, p, z1, aux, params, grid, n); pp = z1; }    /* :71 */
        ax_op(pp, ap, grid);                                                  /* :31 / :72 */
        double rr0 = ko_dot(r, r0, n);                                        /* :32 */
        double alpha = rr0 / ko_dot(ap, r0, n);                               /* :33 */
        for (int64_t t = 0; t < n; ++t) s[t] = fma(-alpha, ap[t], r[t]);      /* :34 */
        if (m_inv) { m_inv(ax_op, s, z2, aux, params, grid, n); ss = z2; }    /* :76 */
        ax_op(ss, as, grid);                                                  /* :35 / :77 */
        double omega = ko_dot(as, s, n) / ko_dot(as, as, n);                  /* :36 */
        for (int64_t t = 0; t < n; ++t)                                       /* :37 / :79 */
            x[t] = fma(omega, ss[t], fma(alpha, pp[t], x[t]));
        for (int64_t t = 0; t < n; ++t) r[t] = fma(-omega, as[t], s[t]);      /* :38 */
        res = ko_norm2(r, n);                                                 /* :39 */
        if (history && hl < history_cap) history[hl] = res;
        ++hl;
        if (res < tol) { *iter = i; break; }                                  /* :40-43 */
        double beta = (ko_dot(r, r0, n) / rr0) * (alpha / omega);             /* :44 */
        for (int64_t t = 0; t < n; ++t)                                       /* :45 */
            p[t] = fma(beta, fma(-omega, ap[t], p[t]), r[t]);
    }
    *res_p = res;
    if (history_len) *history_len = hl;
    free(r); free(r0); free(ap); free(s); free(as); free(p); free(z1); free(z2); free(aux);
    return 0;
}

/* bicgstab.f90:91-182 pbicgstab_omp.  The accumulators rr0, ap_r0, as_s,
 * as_as, r_r0_new are never initialised in the reference before iteration 1
 * (:102) -- treated as 0 (what gfortran's stack gives in practice and the only
 * value for which the routine is correct).  `iters` is undefined when the
 * solver does not converge (:157,181): here *max_iter is left unchanged.
 * m_inv == NULL runs the same loop unpreconditioned (z1 = p, z2 = s). */
int ko_pbicgstab_omp(ko_stencil_fn ax_op, const double *b, int64_t n, double *x, double tol,
                     int *max_iter, double *res_p, ko_precond_fn m_inv, const double *params,
                     double *history, int history_cap, int *history_len)
{
    int grid = ko_grid_side(n), hl = 0;
    double *r = malloc(sizeof(double) * n), *r0 = malloc(sizeof(double) * n);
    double *ap = malloc(sizeof(double) * n), *s = malloc(sizeof(double) * n);
    double *as = malloc(sizeof(double) * n), *p = malloc(sizeof(double) * n);
    double *z1 = malloc(sizeof(double) * n), *z2 = malloc(sizeof(double) * n);
    double *aux = malloc(sizeof(double) * n);
    int converged = 0, iters = *max_iter, maxit = *max_iter;
    double rr0 = 0.0, ap_r0 = 0.0, as_s = 0.0, as_as = 0.0, r_r0_new = 0.0;
    double alpha = 0.0, beta = 0.0, omega = 0.0, res = 0.0;
#pragma omp parallel
    {
#pragma omp for
        for (int64_t t = 0; t < n; ++t) { x[t] = 0.0; r[t] = b[t]; r0[t] = r[t]; p[t] = r0[t]; } /* :114-118 */
        for (int i = 1; i <= maxit; ++i) {
            if (converged) continue;                                          /* :120 */
            if (m_inv) m_inv(ax_op, p, z1, aux, params, grid, n);             /* :121 */
            else {
#pragma omp for
                for (int64_t t = 0; t < n; ++t) z1[t] = p[t];
            }
            ax_op(z1, ap, grid);                                              /* :122 */
#pragma omp for reduction(+ : rr0, ap_r0)
            for (int64_t t = 0; t < n; ++t) {                                 /* :123-128 */
                rr0 = fma(r[t], r0[t], rr0);
                ap_r0 = fma(ap[t], r0[t], ap_r0);
            }
#pragma omp single
            alpha = rr0 / ap_r0;                                              /* :129-131 */
#pragma omp for
            for (int64_t t = 0; t < n; ++t) s[t] = fma(-alpha, ap[t], r[t]);  /* :132-136 */
            if (m_inv) m_inv(ax_op, s, z2, aux, params, grid, n);             /* :137 */
            else {
#pragma omp for
                for (int64_t t = 0; t < n; ++t) z2[t] = s[t];
            }
            ax_op(z2, as, grid);                                              /* :138 */
#pragma omp for reduction(+ : as_s, as_as)
            for (int64_t t = 0; t < n; ++t) {                                 /* :139-144 */
                as_s = fma(as[t], s[t], as_s);
                as_as = fma(as[t], as[t], as_as);
            }
#pragma omp single
            omega = as_s / as_as;                                             /* :145-147 */
#pragma omp for
            for (int64_t t = 0; t < n; ++t) {                                 /* :148-153 */
                x[t] = fma(omega, z2[t], fma(alpha, z1[t], x[t]));
                r[t] = fma(-omega, as[t], s[t]);
            }
#pragma omp single
            {
                res = ko_norm2(r, n);                                         /* :154-160 */
                if (history && hl < history_cap) history[hl] = res;
                ++hl;
                if (res < tol) { iters = i; converged = 1; }
            }
#pragma omp for reduction(+ : r_r0_new)
            for (int64_t t = 0; t < n; ++t) r_r0_new = fma(r[t], r0[t], r_r0_new); /* :161-165 */
#pragma omp single
            {
                beta = (r_r0_new / rr0) * (alpha / omega);                    /* :166-173 */
                r_r0_new = 0.0; as_s = 0.0; as_as = 0.0; rr0 = 0.0; ap_r0 = 0.0;
            }
#pragma omp for
            for (int64_t t = 0; t < n; ++t)                                   /* :174-178 */
                p[t] = fma(beta, fma(-omega, ap[t], p[t]), r[t]);
        }
    }
    *max_iter = iters;                                                        /* :181 */
    *res_p = res;
    if (history_len) *history_len = hl;
    free(r); free(r0); free(ap); free(s); free(as); free(p); free(z1); free(z2); free(aux);
    return 0;
}

/* --------------------------------------------------------------------- */
/* driver-side helpers (tests/test_poisson_mf.f90:39-40 etc.)            */
/* --------------------------------------------------------------------- */

/* x = 1; b = A*x, called outside any parallel region => serial. */
void ko_manufactured_rhs(ko_stencil_fn A, int nsize, double *b)
{
    int64_t n = (int64_t)nsize * nsize;
    double *x = malloc(sizeof(double) * n);
    for (int64_t t = 0; t < n; ++t) x[t] = 1.0;
    A(x, b, nsize);
    free(x);
}

/* run an operator / preconditioner as a team (for timing) or serially */
void ko_apply(ko_stencil_fn A, const double *x, double *y, int nsize, int parallel)
{
    if (parallel) {
#pragma omp parallel
        A(x, y, nsize);
    } else {
        A(x, y, nsize);
    }
}
void ko_apply_precond(ko_precond_fn M, ko_stencil_fn A, const double *r, double *z, double *aux,
                      const double *params, int nsize, int64_t len, int parallel)
{
    if (parallel) {
#pragma omp parallel
        M(A, r, z, aux, params, nsize, len);
    } else {
        M(A, r, z, aux, params, nsize, len);
    }
}

/* function-pointer getters so that ctypes callers can pass the built-ins */
ko_stencil_fn ko_get_stvec(void) { return ko_stvec; }
ko_stencil_fn ko_get_stv_poisson(void) { return ko_stv_poisson; }
ko_precond_fn ko_get_cbpr2(void) { return ko_cbpr2; }
ko_precond_fn ko_get_identity(void) { return ko_precond_identity; }
