// kl_ops.cuh -- problem descriptor, halo exchange, operator / preconditioner application.
#pragma once
#include "kl_functors.cuh"
#include "kl_chain_tma.cuh"

namespace kl {

struct Prob {
    Ctx *c;
    kl_operator_t op;
    kl_precond_t pc;
    double params[8];
    int nparams;
    int nx, ny;        // global grid
    int nyl, j0;       // this rank's slab
    size_t n;          // nx * nyl
    bool builtin_op() const {   // the stencil operators (fused kernels); user callbacks and dense matrices take the generic path
        return op.kind == KL_OP_POISSON5 || op.kind == KL_OP_POISSON5_BRANCHY || op.kind == KL_OP_ANISO5;
    }
};

int prob_init(Prob *P, Ctx *c, const kl_operator_t *op, const kl_precond_t *pc, const double *params,
              int nparams, int nx, int ny);

// Halo buffers: slot s in [0,4) -> lo/hi line of nx doubles each.
struct Halo {
    const double *lo[4];
    const double *hi[4];
};
// exchange the boundary lines of `nvec` vectors; fills H (nullptr at the global
// boundary / single GPU).
int halo_exchange(Prob *P, const double *const *vecs, int nvec, Halo *H);

int halo_exchange_lines(Prob *P, const double *const *vecs, int nvec, int nlines, Halo *H);
// can the temporally blocked kernels run L levels on this problem?  (all ranks take the same decision)
inline bool chain_ok(const Prob *P, int L) {
    const Ctx *c = P->c;
    if (!(c->opt_tma && c->opt_fuse && c->opt_chain) || P->nx % 2 != 0 || P->nx < 64) return false;
    if (c->nranks == 1) return true;
    return P->ny / c->nranks >= L && (size_t)L * P->nx <= 65536;   // L-line halo fits the exchange buffers
}

template <class F>
inline void set_io(F &f, const Prob *P, const double *const *vecs, const Halo &H) {
    for (int a = 0; a < F::NIN; ++a) {
        f.in[a] = vecs[a];
        f.lo[a] = H.lo[a];
        f.hi[a] = H.hi[a];
    }
}
template <class F>
inline void set_gate(F &f, const Ctx *c, bool gated, int step = 0, int run_on_conv = 0) {
    f.flags = gated ? c->d_I : nullptr;
    f.step = step;
    f.run_on_conv = run_on_conv;
}

// y = A x
int op_apply(Prob *P, const double *x, double *y, bool gated);
// y = matmul(A, x) for a dense column-major n x n matrix in device memory (kl_dense.cu)
int launch_gemv(Ctx *c, const double *dA, int n, const double *x, double *y, bool gated);
// solver bodies shared by the stencil and the dense entry points
int gmres_mgsr_solve(Ctx *c, const kl_operator_t *A, const double *b, double *x, int nx, int ny, int m, double tol,
                     double *final_err, double *v_err, int *n_out_p, int *restart_out_p, const kl_precond_t *M,
                     const double *params, int nparams, int mf);
int gmres_hh_solve(Ctx *c, const kl_operator_t *A, const double *b, double *x, int nx, int ny, int m, double tol,
                   double *final_err, double *v_err, int *n_out_p, int *stages_out_p, const kl_precond_t *M,
                   const double *params, int nparams, int prec_variant);
// z = b - A x
int op_resid(Prob *P, const double *x, const double *b, double *z, bool gated);
// z = M^-1 r ; mode 0: no reduction, 1: S_RED[0] = sum z*z, 2: S_RED[0] = sum r*z.
// aux/aux2: scratch vectors (aux2 only needed for KL_PC_CHEB).  For KL_PC_NONE
// z = r is a copy.  `post` runs after the reduction (mode != 0).
template <class Post>
int pc_apply(Prob *P, const double *r, double *z, double *aux, double *aux2, int mode, bool gated,
             const Post &post);

// staging of user vectors according to the pointer mode
int stage_in(Ctx *c, double *d_dst, const double *src, size_t n);
int stage_out(Ctx *c, double *dst, const double *d_src, size_t n);

// read the scalar / int blocks back (async copy + sync)
int read_back(Ctx *c);
// copy history from the device and fill ctx->history
int fetch_history(Ctx *c);

struct PostStoreRed {   // S[dst] = S_RED[0]  (optionally sqrt)
    double *S;
    int dst;
    int do_sqrt;
    __device__ __forceinline__ void run() const {
        double v = S[S_RED];
        S[dst] = do_sqrt ? sqrt(v) : v;
    }
};

// Chebyshev degree-k step (see oracle/krylov_extras.c ko_cheb):
//   first: u = r/theta (= d_old) ; else u = z_old (in[0]), d_old from memory
//   d_new = fma(c1, d_old, c2*(r - A u)) ; z_new = u + d_new
template <int MODE>
struct FChebStep : StencilBase<1, (MODE ? 1 : 0)> {
    const double *r;
    double *d, *z_new;
    double theta, c1, c2;
    int first;
    FastDiv fd;
    __device__ __forceinline__ void init() { fd.set(theta); }
    __device__ __forceinline__ double point(const double (&v)[1]) const { return first ? fd.div(v[0]) : v[0]; }
    template <int VEC>
    __device__ __forceinline__ void store(size_t idx, const double (&raw)[1][VEC], const double (&cu)[VEC],
                                          const double (&au)[VEC], double *acc) const {
        double rr[VEC], dd[VEC], zz[VEC];
        if (first) {
#pragma unroll
            for (int v = 0; v < VEC; ++v) { rr[v] = raw[0][v]; dd[v] = cu[v]; }
        } else {
            KL_LD(VEC, rr, r, idx)
            if (VEC == 2) {
                double2 t = *reinterpret_cast<const double2 *>(d + idx);
                dd[0] = t.x; dd[VEC - 1] = t.y;
            } else {
                dd[0] = d[idx];
            }
        }
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            dd[v] = fma(c1, dd[v], c2 * (rr[v] - au[v]));
            zz[v] = cu[v] + dd[v];
            if (MODE == 1) acc[0] = fma(zz[v], zz[v], acc[0]);
            if (MODE == 2) acc[0] = fma(rr[v], zz[v], acc[0]);
        }
        KL_ST(VEC, d, idx, dd)
        KL_ST(VEC, z_new, idx, zz)
    }
};

template <class Post>
int pc_apply(Prob *P, const double *r, double *z, double *aux, double *aux2, int mode, bool gated,
             const Post &post) {
    Ctx *c = P->c;
    const int kind = P->pc.kind;
    if (kind == KL_PC_CBPR2 && P->builtin_op()) {
        Cbpr2Coef cf = cbpr2_coef(P->params);
        Halo H;
        const double *vecs[1] = {r};
        KL_TRY(halo_exchange(P, vecs, 1, &H));
#define KL_CB(MODE)                                     \
    {                                                   \
        FCbpr2<MODE> f;                                 \
        set_io(f, P, vecs, H);                          \
        set_gate(f, c, gated);                          \
        f.z = z;                                        \
        f.d = cf.d;                                     \
        f.alpha = cf.alpha;                             \
        return launch_stencil(c, &P->op, f, P->nx, P->nyl, post); \
    }
        if (mode == 0) KL_CB(0) else if (mode == 1) KL_CB(1) else KL_CB(2)
#undef KL_CB
    }
    if (kind == KL_PC_CHEB && P->builtin_op() && P->pc.degree >= 1) {
        // Saad Alg. 12.1; coefficients on the host (same arithmetic as the oracle)
        double ea = P->params[0], eb = P->params[1];
        double theta = (eb + ea) / 2.0, delta = fabs(eb - ea) / 2.0;
        double sigma = theta / delta, rho_prev = 1.0 / sigma;
        const int k = P->pc.degree;
        int s0 = 0;
        // ping-pong so that the final result lands in z: z_k = z if k odd else aux
        double *zb[2] = {z, aux};
        // degree <= 6: one chain ; 7..12: two balanced chains (the second continues from the stored z and d) ;
        // above: a chain of 6 and one pass per remaining step
        const bool two = k > kChainMaxL && k <= 2 * kChainMaxL;
        const int kc = two ? (k + 1) / 2 : (k < kChainMaxL ? k : kChainMaxL);
        if (chain_ok(P, kc)) {
            // the first min(k, 6) steps in ONE pass over r (kl_chain_tma.cuh): 16n B instead of 40n B per step
            Halo H;
            const double *hv[1] = {r};
            KL_TRY(halo_exchange_lines(P, hv, 1, kc, &H));
            double c1s[kChainMaxL], c2s[kChainMaxL];
            for (int s = 0; s < kc; ++s) {
                double rho = 1.0 / (2.0 * sigma - rho_prev);
                c1s[s] = rho * rho_prev;
                c2s[s] = 2.0 * rho / delta;
                rho_prev = rho;
            }
            double *dst = (kc == k) ? z : (two ? aux : zb[(k - kc) & 1]);
#define KL_CC(LL)                                                                       \
    case LL: {                                                                          \
        ChCheb<LL> f;                                                                   \
        set_io(f, P, hv, H);                                                            \
        set_gate(f, c, gated);                                                          \
        f.z = dst; f.d_out = (kc == k) ? nullptr : aux2; f.mode = (kc == k) ? mode : 0; \
        f.theta = theta;                                                                \
        for (int s = 0; s < LL; ++s) { f.c1[s] = c1s[s]; f.c2[s] = c2s[s]; }            \
        if (kc == k && mode != 0) { KL_TRY(launch_chain(c, &P->op, f, P->nx, P->nyl, post)); } \
        else { KL_TRY(launch_chain(c, &P->op, f, P->nx, P->nyl, NoPost{})); }           \
    } break;
            switch (kc) {
                KL_CC(1) KL_CC(2) KL_CC(3) KL_CC(4) KL_CC(5) KL_CC(6)
                default: return c->fail(KL_ERR_INVALID, "chebyshev chain length");
            }
#undef KL_CC
            if (kc == k) return KL_OK;
            if (two) {
                // second chain: (z_kc, d_kc, r) -> z_k.  Inputs aux, aux2 ; output z.
                const int kb = k - kc;
                Halo H2;
                const double *hv3[3] = {aux, aux2, r};
                KL_TRY(halo_exchange_lines(P, hv3, 3, kb, &H2));
                for (int s = 0; s < kb; ++s) {
                    double rho = 1.0 / (2.0 * sigma - rho_prev);
                    c1s[s] = rho * rho_prev;
                    c2s[s] = 2.0 * rho / delta;
                    rho_prev = rho;
                }
#define KL_CT(LL)                                                                       \
    case LL: {                                                                          \
        ChChebCont<LL> f;                                                               \
        set_io(f, P, hv3, H2);                                                          \
        set_gate(f, c, gated);                                                          \
        f.z = z; f.mode = mode;                                                         \
        for (int s = 0; s < LL; ++s) { f.c1[s] = c1s[s]; f.c2[s] = c2s[s]; }            \
        if (mode != 0) { KL_TRY(launch_chain(c, &P->op, f, P->nx, P->nyl, post)); }     \
        else { KL_TRY(launch_chain(c, &P->op, f, P->nx, P->nyl, NoPost{})); }           \
    } break;
                switch (kb) {
                    KL_CT(1) KL_CT(2) KL_CT(3) KL_CT(4) KL_CT(5) KL_CT(6)
                    default: return c->fail(KL_ERR_INVALID, "chebyshev chain length");
                }
#undef KL_CT
                return KL_OK;
            }
            s0 = kc;
        }
        for (int s = s0; s < k; ++s) {
            double rho = 1.0 / (2.0 * sigma - rho_prev);
            double c1 = rho * rho_prev, c2 = 2.0 * rho / delta;
            const double *src = (s == 0) ? r : zb[(k - s) & 1];
            double *dst = zb[(k - s - 1) & 1];
            Halo H;
            const double *vecs[1] = {src};
            KL_TRY(halo_exchange(P, vecs, 1, &H));
            const int m = (s == k - 1) ? mode : 0;
#define KL_CH(MODE)                                                       \
    {                                                                     \
        FChebStep<MODE> f;                                                \
        set_io(f, P, vecs, H);                                            \
        set_gate(f, c, gated);                                            \
        f.r = r; f.d = aux2; f.z_new = dst;                               \
        f.theta = theta; f.c1 = c1; f.c2 = c2; f.first = (s == 0);        \
        if (MODE == 0) { KL_TRY(launch_stencil(c, &P->op, f, P->nx, P->nyl, NoPost{})); } \
        else { KL_TRY(launch_stencil(c, &P->op, f, P->nx, P->nyl, post)); } \
    }
            if (m == 0) KL_CH(0) else if (m == 1) KL_CH(1) else KL_CH(2)
#undef KL_CH
            rho_prev = rho;
        }
        return KL_OK;
    }
    // generic path: KL_PC_NONE, user preconditioner, or built-in preconditioner on a user operator
    if (kind == KL_PC_NONE) {
        if (z != r) {
            // gated copy
            PCopy f;
            set_gate(f, c, gated);
            f.a = r; f.y = z;
            KL_TRY(launch_pointwise(c, f, P->n, NoPost{}));
        }
    } else if (kind == KL_PC_USER) {
        if (!P->pc.fn) return c->fail(KL_ERR_INVALID, "KL_PC_USER without callback");
        int rc = P->pc.fn(c, &P->op, P->pc.user, r, z, aux, P->params, P->nparams, P->nx, P->nyl,
                          (void *)c->stream);
        if (rc != 0) return c->fail(KL_ERR_INVALID, "user preconditioner failed");
    } else if (kind == KL_PC_CBPR2) {
        // user operator: the reference's three loops (chebyshev.f90:27-37)
        Cbpr2Coef cf = cbpr2_coef(P->params);
        cudaMemcpyAsync(c->d_S + S_CD, &cf.d, sizeof(double), cudaMemcpyHostToDevice, c->stream);
        PScale s;
        set_gate(s, c, gated);
        s.in = r; s.out = z; s.S = c->d_S; s.s_idx = S_CD;
        KL_TRY(launch_pointwise(c, s, P->n, NoPost{}));
        KL_TRY(op_apply(P, z, aux, gated));
        // z = z + alpha*(r - aux): t = r - aux (in aux), then z += alpha*t
        cudaMemcpyAsync(c->d_S + S_CALPHA, &cf.alpha, sizeof(double), cudaMemcpyHostToDevice, c->stream);
        double one = 1.0;
        cudaMemcpyAsync(c->d_S + S_TMP3, &one, sizeof(double), cudaMemcpyHostToDevice, c->stream);
        PAxpy t;
        set_gate(t, c, gated);
        t.a = r; t.b = aux; t.y = aux; t.S = c->d_S; t.s_idx = S_TMP3; t.sign = -1.0;
        KL_TRY(launch_pointwise(c, t, P->n, NoPost{}));
        PAxpy u;
        set_gate(u, c, gated);
        u.a = z; u.b = aux; u.y = z; u.S = c->d_S; u.s_idx = S_CALPHA; u.sign = 1.0;
        KL_TRY(launch_pointwise(c, u, P->n, NoPost{}));
    } else {
        return c->fail(KL_ERR_UNSUPPORTED, "preconditioner kind not supported with this operator");
    }
    if (mode != 0) {
        PDot2 dt;
        set_gate(dt, c, gated);
        dt.a = (mode == 1) ? z : r; dt.b = z; dt.c = nullptr; dt.d = nullptr;
        // PDot2 has NRED = 2; post reads S_RED[0] only
        KL_TRY(launch_pointwise(c, dt, P->n, post));
    }
    return KL_OK;
}

}  // namespace kl
