// kl_ops.cu -- operator / preconditioner application and the public apply entry points.
#include <math.h>
#include <string.h>

#include "kl_ops.cuh"

namespace kl {

int prob_init(Prob *P, Ctx *c, const kl_operator_t *op, const kl_precond_t *pc, const double *params,
              int nparams, int nx, int ny) {
    if (!c || !op) return KL_ERR_INVALID;
    if (op->kind == KL_OP_DENSE) {
        if (nx < 2 || ny != 1) return c->fail(KL_ERR_INVALID, "dense operator: nx = n >= 2, ny = 1");
        if (!op->user) return c->fail(KL_ERR_INVALID, "KL_OP_DENSE without a matrix");
        if (c->nranks > 1) return c->fail(KL_ERR_UNSUPPORTED, "dense operators are single-GPU only");
    } else if (nx < 2 || ny < 2) return c->fail(KL_ERR_INVALID, "grid must be at least 2x2");
    P->c = c;
    P->op = *op;
    if (pc) P->pc = *pc;
    else P->pc = kl_precond_t{KL_PC_NONE, 0, nullptr, nullptr};
    if (P->op.kind != KL_OP_POISSON5 && P->op.kind != KL_OP_POISSON5_BRANCHY && P->op.kind != KL_OP_ANISO5 &&
        P->op.kind != KL_OP_USER && P->op.kind != KL_OP_DENSE && P->op.kind != KL_OP_ANISO5_VAR)
        return c->fail(KL_ERR_INVALID, "unknown operator kind");
    if (P->op.kind == KL_OP_ANISO5_VAR) {
        const kl_aniso_var_t *cf = (const kl_aniso_var_t *)P->op.user;
        if (!cf || !cf->kx || !cf->ky) return c->fail(KL_ERR_INVALID, "KL_OP_ANISO5_VAR without coefficient arrays");
        if (c->nranks > 1) return c->fail(KL_ERR_UNSUPPORTED, "variable-coefficient operators are single-GPU only");
        if (ny > 65535) return c->fail(KL_ERR_UNSUPPORTED, "KL_OP_ANISO5_VAR: ny <= 65535");
    }
    if (P->op.kind == KL_OP_USER && !P->op.fn) return c->fail(KL_ERR_INVALID, "KL_OP_USER without callback");
    if (P->op.kind == KL_OP_USER && c->nranks > 1)
        return c->fail(KL_ERR_UNSUPPORTED, "user operators are single-GPU only");
    if (P->op.kind != KL_OP_ANISO5) { P->op.eps_x = 1.0; P->op.eps_y = 1.0; }
    P->nparams = nparams < 8 ? nparams : 8;
    for (int i = 0; i < 8; ++i) P->params[i] = (params && i < P->nparams) ? params[i] : 0.0;
    if ((P->pc.kind == KL_PC_CBPR2 || P->pc.kind == KL_PC_CHEB) && P->nparams < 2)
        return c->fail(KL_ERR_INVALID, "Chebyshev preconditioners need params(1:2)");
    P->nx = nx;
    P->ny = ny;
    kl_partition(c, ny, &P->j0, &P->nyl);
    if (P->nyl < 1) return c->fail(KL_ERR_INVALID, "more ranks than grid lines");
    P->n = (size_t)nx * P->nyl;
    KL_CUDA(c, cudaSetDevice(c->device));
    if (c->nranks > 1) {
        size_t need = (size_t)nx * 8 * kChainMaxL;  // 4 slots x {lo,hi} x up to kChainMaxL lines (chained kernels)
        if (c->halo_doubles < need) {
            cudaStreamSynchronize(c->stream);
            cudaFree(c->d_halo);
            c->d_halo = nullptr;
            cudaError_t e = cudaMalloc(&c->d_halo, need * sizeof(double));
            if (e != cudaSuccess) return c->fail(KL_ERR_ALLOC, "halo buffers", e);
            c->halo_doubles = need;
        }
    }
    return KL_OK;
}

int halo_exchange(Prob *P, const double *const *vecs, int nvec, Halo *H) {
    Ctx *c = P->c;
    for (int a = 0; a < 4; ++a) H->lo[a] = H->hi[a] = nullptr;
    if (c->nranks == 1) return KL_OK;
    if (nvec > 4) return c->fail(KL_ERR_INVALID, "halo_exchange: too many vectors");
    const double *slo[4], *shi[4];
    for (int a = 0; a < nvec; ++a) {
        slo[a] = vecs[a];
        shi[a] = vecs[a] + (size_t)(P->nyl - 1) * P->nx;
    }
    // NOTE: a stencil kernel treats lo[0]==nullptr as "no lower neighbour" for all inputs.
    return comm_halo_exchange(c, slo, shi, H->lo, H->hi, nvec, P->nx);
}

// L-line halo for the temporally blocked kernels: lo = the lower neighbour's last `nlines` lines (grid lines
// -nlines..-1 in order), hi = the upper neighbour's first `nlines` lines.  Needs nyl >= nlines on every rank.
int halo_exchange_lines(Prob *P, const double *const *vecs, int nvec, int nlines, Halo *H) {
    Ctx *c = P->c;
    for (int a = 0; a < 4; ++a) H->lo[a] = H->hi[a] = nullptr;
    if (c->nranks == 1) return KL_OK;
    if (nvec > 4 || nlines < 1 || nlines > P->nyl) return c->fail(KL_ERR_INVALID, "halo_exchange_lines: bad arguments");
    const double *slo[4], *shi[4];
    for (int a = 0; a < nvec; ++a) {
        slo[a] = vecs[a];
        shi[a] = vecs[a] + (size_t)(P->nyl - nlines) * P->nx;
    }
    return comm_halo_exchange(c, slo, shi, H->lo, H->hi, nvec, nlines * P->nx);
}

// Variable-coefficient anisotropic diffusion, y = A(kx, ky) x (definition and evaluation order: oracle/krylov_extras.c
// ko_aniso_var).  One thread per pair of grid points, neighbours through L1/L2; the operator takes the generic solver
// path (one kernel per reference loop), like a user operator.  32n B (x, kx, ky in, y out).
__global__ void __launch_bounds__(256)
k_aniso_var(const double *__restrict__ x, double *__restrict__ y, const double *__restrict__ kx,
            const double *__restrict__ ky, const int nx, const int ny, const int *__restrict__ flags) {
    if (flags && flags[I_CONV_AT] >= 0) return;
    const int i = blockIdx.x * 256 + threadIdx.x, j = blockIdx.y;
    if (i >= nx) return;
    const size_t idx = (size_t)i + (size_t)j * nx;
    const double kxc = __ldg(kx + idx), kyc = __ldg(ky + idx), xc = __ldg(x + idx);
    const bool hl = i > 0, hr = i < nx - 1, hu = j > 0, hd = j < ny - 1;
    const double wl = hl ? 0.5 * (kxc + __ldg(kx + idx - 1)) : kxc, wr = hr ? 0.5 * (kxc + __ldg(kx + idx + 1)) : kxc;
    const double wu = hu ? 0.5 * (kyc + __ldg(ky + idx - nx)) : kyc, wd = hd ? 0.5 * (kyc + __ldg(ky + idx + nx)) : kyc;
    const double xl = hl ? __ldg(x + idx - 1) : 0.0, xr = hr ? __ldg(x + idx + 1) : 0.0;
    const double xd = hd ? __ldg(x + idx + nx) : 0.0, xu = hu ? __ldg(x + idx - nx) : 0.0;
    const double diag = ((wl + wr) + wd) + wu;
    const double s = fma(wl, xl, wr * xr), t = fma(wd, xd, wu * xu);
    y[idx] = fma(diag, xc, -(s + t));
}

int op_apply(Prob *P, const double *x, double *y, bool gated) {
    Ctx *c = P->c;
    if (P->op.kind == KL_OP_ANISO5_VAR) {
        const kl_aniso_var_t *cf = (const kl_aniso_var_t *)P->op.user;
        dim3 grid((unsigned)((P->nx + 255) / 256), (unsigned)P->nyl);
        k_aniso_var<<<grid, 256, 0, c->stream>>>(x, y, cf->kx, cf->ky, P->nx, P->nyl, gated ? c->d_I : nullptr);
        c->stats.kernel_launches++;
        return KL_OK;
    }
    if (P->op.kind == KL_OP_DENSE) return launch_gemv(c, (const double *)P->op.user, P->nx, x, y, gated);
    if (!P->builtin_op()) {
        // user callbacks cannot be gated on the device flag; they run unconditionally
        int rc = P->op.fn(P->op.user, x, y, P->nx, P->nyl, (void *)c->stream);
        if (rc != 0) return c->fail(KL_ERR_INVALID, "user operator failed");
        return KL_OK;
    }
    Halo H;
    const double *vecs[1] = {x};
    KL_TRY(halo_exchange(P, vecs, 1, &H));
    FApply f;
    set_io(f, P, vecs, H);
    set_gate(f, c, gated);
    f.y = y;
    return launch_stencil(c, &P->op, f, P->nx, P->nyl, NoPost{});
}

int op_resid(Prob *P, const double *x, const double *b, double *z, bool gated) {
    Ctx *c = P->c;
    if (!P->builtin_op()) {
        KL_TRY(op_apply(P, x, z, gated));
        double one = 1.0;
        KL_CUDA(c, cudaMemcpyAsync(c->d_S + S_TMP3, &one, sizeof(double), cudaMemcpyHostToDevice, c->stream));
        PAxpy t;
        set_gate(t, c, gated);
        t.a = b; t.b = z; t.y = z; t.S = c->d_S; t.s_idx = S_TMP3; t.sign = -1.0;
        return launch_pointwise(c, t, P->n, NoPost{});
    }
    Halo H;
    const double *vecs[1] = {x};
    KL_TRY(halo_exchange(P, vecs, 1, &H));
    FResid f;
    set_io(f, P, vecs, H);
    set_gate(f, c, gated);
    f.b = b;
    f.z = z;
    return launch_stencil(c, &P->op, f, P->nx, P->nyl, NoPost{});
}

int stage_in(Ctx *c, double *d_dst, const double *src, size_t n) {
    cudaMemcpyKind k = c->pointer_mode == KL_POINTER_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
    if (c->pointer_mode == KL_POINTER_DEVICE && d_dst == src) return KL_OK;
    KL_CUDA(c, cudaMemcpyAsync(d_dst, src, n * sizeof(double), k, c->stream));
    if (k == cudaMemcpyHostToDevice) c->stats.h2d_bytes += (double)n * sizeof(double);
    return KL_OK;
}
int stage_out(Ctx *c, double *dst, const double *d_src, size_t n) {
    cudaMemcpyKind k = c->pointer_mode == KL_POINTER_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
    if (c->pointer_mode == KL_POINTER_DEVICE && dst == d_src) return KL_OK;
    KL_CUDA(c, cudaMemcpyAsync(dst, d_src, n * sizeof(double), k, c->stream));
    if (k == cudaMemcpyDeviceToHost) c->stats.d2h_bytes += (double)n * sizeof(double);
    return KL_OK;
}

int read_back(Ctx *c) {
    KL_CUDA(c, cudaMemcpyAsync(c->h_pinned, c->d_S, sizeof(double) * 64, cudaMemcpyDeviceToHost, c->stream));
    KL_CUDA(c, cudaMemcpyAsync(c->h_pinned_i, c->d_I, sizeof(int) * I_COUNT, cudaMemcpyDeviceToHost, c->stream));
    KL_CUDA(c, cudaStreamSynchronize(c->stream));
    return KL_OK;
}

int fetch_history(Ctx *c) {
    KL_TRY(read_back(c));
    int len = c->h_pinned_i[I_HIST];
    c->history_len = len;
    int k = len < c->hist_cap ? len : c->hist_cap;
    c->history.resize(k);
    if (k > 0) {
        KL_CUDA(c, cudaMemcpyAsync(c->history.data(), c->d_hist, sizeof(double) * k, cudaMemcpyDeviceToHost, c->stream));
        KL_CUDA(c, cudaStreamSynchronize(c->stream));
    }
    return KL_OK;
}

}  // namespace kl

using namespace kl;

extern "C" {

int kl_apply_operator(kl_handle_t h, const kl_operator_t *A_x, const double *x, double *y, int nx, int ny) {
    if (!h || !A_x || !x || !y) return KL_ERR_INVALID;
    Ctx *c = h;
    Prob P;
    KL_TRY(prob_init(&P, c, A_x, nullptr, nullptr, 0, nx, ny));
    if (c->pointer_mode == KL_POINTER_DEVICE) {
        KL_TRY(op_apply(&P, x, y, false));
        return KL_OK;
    }
    KL_TRY(ws_reserve(c, 2 * ws_need(P.n)));
    ws_reset(c);
    double *dx = ws_take<double>(c, P.n), *dy = ws_take<double>(c, P.n);
    KL_TRY(stage_in(c, dx, x, P.n));
    KL_TRY(op_apply(&P, dx, dy, false));
    KL_TRY(stage_out(c, y, dy, P.n));
    KL_CUDA(c, cudaStreamSynchronize(c->stream));
    KL_CUDA(c, cudaGetLastError());
    return KL_OK;
}

int kl_apply_precond(kl_handle_t h, const kl_precond_t *M_inv, const kl_operator_t *A_x, const double *r,
                     double *z, const double *params, int nparams, int nx, int ny) {
    if (!h || !A_x || !M_inv || !r || !z) return KL_ERR_INVALID;
    Ctx *c = h;
    Prob P;
    KL_TRY(prob_init(&P, c, A_x, M_inv, params, nparams, nx, ny));
    KL_TRY(ws_reserve(c, 4 * ws_need(P.n)));
    ws_reset(c);
    if (c->pointer_mode == KL_POINTER_DEVICE && r != z) {
        // device vectors are used in place (stream-ordered, returns after enqueue like kl_apply_operator)
        double *aux = ws_take<double>(c, P.n), *aux2 = ws_take<double>(c, P.n);
        KL_TRY(pc_apply(&P, r, z, aux, aux2, 0, false, NoPost{}));
        KL_CUDA(c, cudaGetLastError());
        return KL_OK;
    }
    double *dr = ws_take<double>(c, P.n), *dz = ws_take<double>(c, P.n);
    double *aux = ws_take<double>(c, P.n), *aux2 = ws_take<double>(c, P.n);
    KL_TRY(stage_in(c, dr, r, P.n));
    KL_TRY(pc_apply(&P, dr, dz, aux, aux2, 0, false, NoPost{}));
    KL_TRY(stage_out(c, z, dz, P.n));
    KL_CUDA(c, cudaStreamSynchronize(c->stream));
    KL_CUDA(c, cudaGetLastError());
    return KL_OK;
}

}  // extern "C"
