// kl_dense.cu -- dense-operator variants of the GMRES solvers and the Hilbert test matrix.
//
// Reference: src/gmres_mgsr.f90:11-95 (gmres_mgsr_dense), src/gmres_hh.f90:10-112 (gmres_hh_dense),
// src/problems/hilbert.f90:6-18 (generate_matrix), tests/test_hilbert.f90.
// The dense solvers are the matrix-free ones with  w = matmul(A, v)  as the operator: the matrix is a
// KL_OP_DENSE plug-in that takes the generic (unfused) path of gmres_mgsr_solve / gmres_hh_solve, so the
// Arnoldi step, the Householder reflectors, the Givens warp and the back-solve are the same device code.
// O(n^2) storage: meant for the reference's small ill-conditioned studies, not for large n.
#include <math.h>

#include "kl_ops.cuh"

namespace kl {

// y = matmul(A, x), A column-major n x n.  One thread per row, j sequential: y(i) is the same left-to-right
// FMA sum as the column sweep y(:) += A(:,j) x(j) (oracle/krylov_oracle.c ko_dense_matvec); consecutive
// threads read consecutive rows of a column (coalesced), x(j) is a broadcast load.
__global__ void __launch_bounds__(128)
k_gemv_cm(const double *__restrict__ A, const int n, const double *__restrict__ x, double *__restrict__ y,
          const int *flags) {
    if (flags && flags[I_CONV_AT] >= 0) return;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double *a = A + i;
    double s0 = 0.0;
    int j = 0;
    for (; j + 4 <= n; j += 4) {
        const double a0 = __ldg(a + (size_t)j * n), a1 = __ldg(a + (size_t)(j + 1) * n);
        const double a2 = __ldg(a + (size_t)(j + 2) * n), a3 = __ldg(a + (size_t)(j + 3) * n);
        s0 = fma(a0, __ldg(x + j), s0);
        s0 = fma(a1, __ldg(x + j + 1), s0);
        s0 = fma(a2, __ldg(x + j + 2), s0);
        s0 = fma(a3, __ldg(x + j + 3), s0);
    }
    for (; j < n; ++j) s0 = fma(__ldg(a + (size_t)j * n), __ldg(x + j), s0);
    y[i] = s0;
}

int launch_gemv(Ctx *c, const double *dA, int n, const double *x, double *y, bool gated) {
    k_gemv_cm<<<(n + 127) / 128, 128, 0, c->stream>>>(dA, n, x, y, gated ? c->d_I : nullptr);
    c->stats.kernel_launches++;
    return KL_OK;
}

// hilbert.f90:13-17  H(i,j) = 1 / real(i+j-1)   (single-precision quotient, then widened)
__global__ void k_hilbert(double *H, const int n) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)n * n) return;
    const int i = (int)(t % n) + 1, j = (int)(t / n) + 1;
    H[t] = (double)__fdiv_rn(1.0f, (float)(i + j - 1));
}

// stage a host matrix into its own device allocation (the workspace arena is reset by the solvers)
struct DenseA {
    Ctx *c;
    const double *d = nullptr;
    double *owned = nullptr;
    int init(Ctx *c_, const double *A, int n) {
        c = c_;
        if (c->pointer_mode == KL_POINTER_DEVICE) { d = A; return KL_OK; }
        cudaError_t e = cudaMalloc(&owned, sizeof(double) * (size_t)n * n);
        if (e != cudaSuccess) return c->fail(KL_ERR_ALLOC, "dense matrix", e);
        KL_CUDA(c, cudaMemcpyAsync(owned, A, sizeof(double) * (size_t)n * n, cudaMemcpyHostToDevice, c->stream));
        c->stats.h2d_bytes += sizeof(double) * (double)n * n;
        d = owned;
        return KL_OK;
    }
    ~DenseA() {
        if (owned) {
            cudaStreamSynchronize(c->stream);
            cudaFree(owned);
        }
    }
};

}  // namespace kl

using namespace kl;

extern "C" {

int kl_gmres_mgsr_dense(kl_handle_t h, const double *A, int n, const double *b, double *x, int m, double tol,
                        double *final_err, double *v_err, int *n_out, int *restart_out) {
    if (!h || !A || n < 2) return KL_ERR_INVALID;
    KL_CUDA(h, cudaSetDevice(h->device));
    DenseA dA;
    KL_TRY(dA.init(h, A, n));
    kl_operator_t op{KL_OP_DENSE, 1.0, 1.0, nullptr, const_cast<double *>(dA.d)};
    // gmres_mgsr.f90:11-95 is the _mf algorithm (:98-199) with matmul as the operator and no preconditioner
    return gmres_mgsr_solve(h, &op, b, x, n, 1, m, tol, final_err, v_err, n_out, restart_out, nullptr, nullptr, 0, 1);
}

int kl_gmres_hh_dense(kl_handle_t h, const double *A, int n, const double *b, double *x, int m, double tol,
                      double *final_err, double *v_err, int *n_out, int *stages_out) {
    if (!h || !A || n < 2) return KL_ERR_INVALID;
    KL_CUDA(h, cudaSetDevice(h->device));
    DenseA dA;
    KL_TRY(dA.init(h, A, n));
    kl_operator_t op{KL_OP_DENSE, 1.0, 1.0, nullptr, const_cast<double *>(dA.d)};
    return gmres_hh_solve(h, &op, b, x, n, 1, m, tol, final_err, v_err, n_out, stages_out, nullptr, nullptr, 0, 2);
}

int kl_generate_matrix(kl_handle_t h, double *H, int n) {
    if (!h || !H || n < 1) return KL_ERR_INVALID;
    Ctx *c = h;
    KL_CUDA(c, cudaSetDevice(c->device));
    const size_t nn = (size_t)n * n;
    if (c->pointer_mode == KL_POINTER_DEVICE) {
        k_hilbert<<<(unsigned)((nn + 255) / 256), 256, 0, c->stream>>>(H, n);
        KL_CUDA(c, cudaGetLastError());
        return KL_OK;
    }
    double *d = nullptr;
    cudaError_t e = cudaMalloc(&d, sizeof(double) * nn);
    if (e != cudaSuccess) return c->fail(KL_ERR_ALLOC, "hilbert matrix", e);
    k_hilbert<<<(unsigned)((nn + 255) / 256), 256, 0, c->stream>>>(d, n);
    cudaMemcpyAsync(H, d, sizeof(double) * nn, cudaMemcpyDeviceToHost, c->stream);
    e = cudaStreamSynchronize(c->stream);
    cudaFree(d);
    if (e != cudaSuccess) return c->fail(KL_ERR_CUDA, "kl_generate_matrix", e);
    return KL_OK;
}

int kl_dense_matvec(kl_handle_t h, const double *A, int n, const double *x, double *y) {
    if (!h || !A || !x || !y || n < 1) return KL_ERR_INVALID;
    Ctx *c = h;
    KL_CUDA(c, cudaSetDevice(c->device));
    DenseA dA;
    KL_TRY(dA.init(c, A, n));
    if (c->pointer_mode == KL_POINTER_DEVICE) {
        KL_TRY(launch_gemv(c, dA.d, n, x, y, false));
        KL_CUDA(c, cudaGetLastError());
        return KL_OK;
    }
    KL_TRY(ws_reserve(c, 2 * ws_need((size_t)n)));
    ws_reset(c);
    double *dx = ws_take<double>(c, n), *dy = ws_take<double>(c, n);
    KL_TRY(stage_in(c, dx, x, n));
    KL_TRY(launch_gemv(c, dA.d, n, dx, dy, false));
    KL_TRY(stage_out(c, y, dy, n));
    KL_CUDA(c, cudaStreamSynchronize(c->stream));
    KL_CUDA(c, cudaGetLastError());
    return KL_OK;
}

}  // extern "C"
