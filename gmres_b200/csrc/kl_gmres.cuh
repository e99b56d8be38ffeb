// kl_gmres.cuh -- pieces shared by the Gram-Schmidt and Householder GMRES solvers.
#pragma once
#include <vector>

#include "kl_ops.cuh"

namespace kl {

constexpr int kTsThreads = 256;
constexpr int kTsWarps = kTsThreads / 32;
constexpr int kTsCpw = 12;          // columns per warp per pass
constexpr int kTsU = 4;             // row chunks per lane per trip
constexpr int kTsMaxBlocks = 1024;  // partials leading dimension


// ---- Givens update by ONE WARP (gmres_mgsr.f90:362-389) --------------------
// hj1 = H(j+1,j) (||w|| after orthogonalisation for MGS, -/+||w(j+1:n)|| for
// Householder).  Called by warp 0 of the last block of the final update kernel
// (single GPU) or by k_givens (multi GPU).  load_h: stage H(0..j,j) from global.
__device__ __forceinline__ void givens_update_warp(const GmresDev &G, int j, double hj1_in, int lane,
                                                   double *sm /* 3*(m+2) doubles */, bool load_h = true) {
    double *sh = sm, *sc = sm + (G.m + 2), *ss = sm + 2 * (G.m + 2);
    double *Hj = G.H + (size_t)j * G.ldh;
    if (load_h)
        for (int i = lane; i <= j; i += 32) sh[i] = Hj[i];
    for (int i = lane; i < j; i += 32) {
        sc[i] = G.cs[i];
        ss[i] = G.sn[i];
    }
    __syncwarp();
    if (lane == 0) {
        const double h_val = fabs(hj1_in);                // :362 norm2(w) / gmres_hh.f90:314
        double hi = sh[0];
        for (int i = 0; i < j; ++i) {                     // :365-369
            const double hn = sh[i + 1], c = sc[i], s = ss[i];
            Hj[i] = fma(c, hi, s * hn);
            hi = fma(-s, hi, c * hn);
        }
        // hi = H(j,j) after the previous rotations; H(j+1,j) = h_val (:363)
        const double hjj = hi, hj1 = hj1_in;
        double ds = hypot(hj1, hjj);                      // :370
        double c = hjj / ds, s = hj1 / ds;                // :371-372
        G.cs[j] = c;
        G.sn[j] = s;
        Hj[j] = fma(c, hjj, s * hj1);                     // :373
        Hj[j + 1] = 0.0;                                  // :374
        double tmp = G.g[j], gn = G.g[j + 1];             // :378-380
        G.g[j] = fma(c, tmp, s * gn);
        double gj1 = fma(-s, tmp, c * gn);
        G.g[j + 1] = gj1;
        double fe = fabs(gj1) / G.S[S_BETA0];             // :383
        G.fe[j] = fe;
        G.S[S_HVAL] = h_val;
        G.S[S_RES] = fe;
        int hl = G.I[I_HIST];
        if (hl < G.hist_cap) G.hist[hl] = fe;
        G.I[I_HIST] = hl + 1;
        G.I[I_ITER] = G.I[I_ITER] + 1;
        G.I[I_NOUT] = j + 1;                              // :389
        const double tol = G.S[S_TOL];
        bool conv = G.mf ? (h_val < tol || fe < tol) : (fe < tol);   // :172 / :385
        if (!(fe == fe)) { G.I[I_BREAKDOWN] = 1; conv = true; }
        if (conv) G.I[I_CONV_AT] = j;
    }
}

// ---- faithful MGS step (gmres_mgsr.f90:342-359), two loops fused in one pass:
//   w -= h_prev * V_prev   (skipped when V_prev == nullptr)
//   acc = V_cur . w        (skipped when V_cur == nullptr; then acc = ||w||^2 if want_norm)
// post (last block): h = acc ; H(i_cur, j) += h ; S_TMP0 = h
struct PMgsStep : PwBase<1> {
    double *w;
    const double *vprev, *vcur;
    const double *S;
    double hprev;
    int want_norm;
    __device__ __forceinline__ void init() { hprev = vprev ? S[S_TMP0] : 0.0; }
    template <int VEC>
    __device__ __forceinline__ void elem(size_t i, double *acc) const {
        double a[VEC];
        if (VEC == 2) {
            double2 t = *reinterpret_cast<const double2 *>(w + i);
            a[0] = t.x; a[VEC - 1] = t.y;
        } else {
            a[0] = w[i];
        }
        if (vprev) {
            double p[VEC];
            KL_LD(VEC, p, vprev, i)
#pragma unroll
            for (int e = 0; e < VEC; ++e) a[e] = fma(-hprev, p[e], a[e]);
            KL_ST(VEC, w, i, a)
        }
        if (vcur) {
            double q[VEC];
            KL_LD(VEC, q, vcur, i)
#pragma unroll
            for (int e = 0; e < VEC; ++e) acc[0] = fma(a[e], q[e], acc[0]);
        } else if (want_norm) {
#pragma unroll
            for (int e = 0; e < VEC; ++e) acc[0] = fma(a[e], a[e], acc[0]);
        }
    }
};

// host launchers implemented in kl_gmres.cu
int launch_vtw(Ctx *c, const double *V, size_t ldv, const double *w, size_t n, int ncols, double *out,
               const GmresDev &G, int j, int h_mode, bool gated);
int launch_backsolve(Ctx *c, const GmresDev &G);
// TMA-staged projection (update = false) or update + second projection (update = true)
bool ts_tma_ok(Ctx *c, size_t n, size_t ldv, int nc, int nx = 0, int ny = 0);
int launch_ts_tma(Ctx *c, bool update, const double *V, size_t ldv, int ncols_total, double *w, size_t n, int nc,
                  const double *h_in, double *out, const GmresDev &G, int j, int h_mode, bool gated,
                  long long tail0 = -1, const double *tail_T = nullptr, double *tail_tvec = nullptr, int tail_ldt = 0);
int gram_lower(Ctx *c, const double *V, size_t ldv, size_t n, int k, double *d_gram, std::vector<double> &out);

}  // namespace kl
