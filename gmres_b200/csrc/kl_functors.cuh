// kl_functors.cuh -- stencil-type and point-wise functors shared by the solvers.
#pragma once
#include "kl_internal.cuh"

namespace kl {

// ---- cbpr2 coefficients, chebyshev.f90:19-25 (host, same arithmetic) -------
struct Cbpr2Coef {
    double d, alpha;
};
inline Cbpr2Coef cbpr2_coef(const double *params) {
    double eigen_min = params[0], eigen_max = params[1];
    double c = (eigen_max - eigen_min) / 2.0;
    double d = (eigen_max + eigen_min) / 2.0;
    double alpha = 1.0 / d;
    double beta = (c * alpha / 2.0) * (c * alpha / 2.0);
    alpha = 1.0 / (d - beta);
    return Cbpr2Coef{d, alpha};
}

// ---------------------------------------------------------------------------
// stencil functors (contract: kl_internal.cuh, k_stencil)
// ---------------------------------------------------------------------------
#define KL_ST(VEC, ptr, idx, src)                            \
    if (VEC == 2) stg2((ptr) + (idx), src[0], src[VEC - 1]); \
    else (ptr)[idx] = src[0];
#define KL_LD(VEC, dst, ptr, idx)                         \
    if (VEC == 2) {                                       \
        double2 t__ = ldg2((ptr) + (idx));                \
        dst[0] = t__.x;                                   \
        dst[VEC - 1] = t__.y;                             \
    } else {                                              \
        dst[0] = __ldg((ptr) + (idx));                    \
    }

// y = A x                                     (poisson.f90:33-77)
struct FApply : StencilBase<1, 0> {
    double *y;
    __device__ __forceinline__ void init() {}
    __device__ __forceinline__ double point(const double (&v)[1]) const { return v[0]; }
    template <int VEC>
    __device__ __forceinline__ void store(size_t idx, const double (&raw)[1][VEC], const double (&cu)[VEC],
                                          const double (&au)[VEC], double *) const {
        KL_ST(VEC, y, idx, au)
    }
};

// y = A x with two fused dot products:
//   acc0 = sum (A x) * e1 ; acc1 = self2 ? sum (A x)^2 : sum (A x) * e2
// (bicgstab.f90:123-127 ap.r0 ; :139-143 as.s, as.as)
struct FApplyDots : StencilBase<1, 2> {
    double *y;
    const double *e1, *e2;
    int self2;
    __device__ __forceinline__ void init() {}
    __device__ __forceinline__ double point(const double (&v)[1]) const { return v[0]; }
    template <int VEC>
    __device__ __forceinline__ void store(size_t idx, const double (&raw)[1][VEC], const double (&cu)[VEC],
                                          const double (&au)[VEC], double *acc) const {
        double a1[VEC], a2[VEC];
        KL_ST(VEC, y, idx, au)
        KL_LD(VEC, a1, e1, idx)
        if (!self2) { KL_LD(VEC, a2, e2, idx) }
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            acc[0] = fma(au[v], a1[v], acc[0]);
            acc[1] = fma(au[v], self2 ? au[v] : a2[v], acc[1]);
        }
    }
};

// z = b - A x                                  (gmres_mgsr.f90:314-319)
struct FResid : StencilBase<1, 0> {
    const double *b;
    double *z;
    __device__ __forceinline__ void init() {}
    __device__ __forceinline__ double point(const double (&v)[1]) const { return v[0]; }
    template <int VEC>
    __device__ __forceinline__ void store(size_t idx, const double (&raw)[1][VEC], const double (&cu)[VEC],
                                          const double (&au)[VEC], double *) const {
        double bb[VEC];
        KL_LD(VEC, bb, b, idx)
#pragma unroll
        for (int v = 0; v < VEC; ++v) bb[v] = bb[v] - au[v];
        KL_ST(VEC, z, idx, bb)
    }
};

// cbpr2 in ONE pass over r (chebyshev.f90:27-37):  u = r/d is point-wise and the
// division is correctly rounded, so A(r/d) can be formed from r directly:
//   z = u + alpha*(r - A u)
// MODE 0: no reduction ; 1: acc0 = sum z*z ; 2: acc0 = sum r*z
template <int MODE>
struct FCbpr2 : StencilBase<1, (MODE ? 1 : 0)> {
    double *z;
    double d, alpha;
    FastDiv fd;
    __device__ __forceinline__ void init() { fd.set(d); }
    __device__ __forceinline__ double point(const double (&v)[1]) const { return fd.div(v[0]); }
    template <int VEC>
    __device__ __forceinline__ void store(size_t idx, const double (&raw)[1][VEC], const double (&cu)[VEC],
                                          const double (&au)[VEC], double *acc) const {
        double zz[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            const double rr = raw[0][v];
            zz[v] = fma(alpha, rr - au[v], cu[v]);
            if (MODE == 1) acc[0] = fma(zz[v], zz[v], acc[0]);
            if (MODE == 2) acc[0] = fma(rr, zz[v], acc[0]);
        }
        KL_ST(VEC, z, idx, zz)
    }
};

// CG direction update fused into the operator (cg.f90:139-142 of the previous
// iteration + :111 + the ax.p half of :118-122):
//   p_new = z + beta*p_old ; ax = A p_new ; acc0 = sum ax*p_new
// in[0] = z (r for plain CG), in[1] = p_old.  p_new goes to a DIFFERENT buffer
// (neighbouring blocks still read p_old).
struct FCgDir : StencilBase<2, 1> {
    double *p_new, *ax;
    const double *S;
    double beta;
    __device__ __forceinline__ void init() { beta = S[S_BETA]; }
    __device__ __forceinline__ double point(const double (&v)[2]) const { return fma(beta, v[1], v[0]); }
    template <int VEC>
    __device__ __forceinline__ void store(size_t idx, const double (&raw)[2][VEC], const double (&cu)[VEC],
                                          const double (&au)[VEC], double *acc) const {
        KL_ST(VEC, p_new, idx, cu)
        KL_ST(VEC, ax, idx, au)
#pragma unroll
        for (int v = 0; v < VEC; ++v) acc[0] = fma(au[v], cu[v], acc[0]);
    }
};

// GMRES: V_{j+1} = w / h_val fused into z = A V_{j+1}
// (gmres_mgsr.f90:384 + :336 of the next step)
struct FScaleApply : StencilBase<1, 0> {
    double *v_out, *z;
    const double *S;
    int s_idx;
    FastDiv fd;
    __device__ __forceinline__ void init() { fd.set(S[s_idx]); }
    __device__ __forceinline__ double point(const double (&v)[1]) const { return fd.div(v[0]); }
    template <int VEC>
    __device__ __forceinline__ void store(size_t idx, const double (&raw)[1][VEC], const double (&cu)[VEC],
                                          const double (&au)[VEC], double *) const {
        KL_ST(VEC, v_out, idx, cu)
        KL_ST(VEC, z, idx, au)
    }
};

// ---------------------------------------------------------------------------
// point-wise functors
// ---------------------------------------------------------------------------
// acc0 = sum a*b ; acc1 = sum c*d   (c == nullptr => only one)
struct PDot2 : PwBase<2> {
    const double *a, *b, *c, *d;
    __device__ __forceinline__ void init() {}
    template <int VEC>
    __device__ __forceinline__ void elem(size_t i, double *acc) const {
        double va[VEC], vb[VEC];
        KL_LD(VEC, va, a, i)
        KL_LD(VEC, vb, b, i)
#pragma unroll
        for (int v = 0; v < VEC; ++v) acc[0] = fma(va[v], vb[v], acc[0]);
        if (c) {
            double vc[VEC], vd[VEC];
            KL_LD(VEC, vc, c, i)
            KL_LD(VEC, vd, d, i)
#pragma unroll
            for (int v = 0; v < VEC; ++v) acc[1] = fma(vc[v], vd[v], acc[1]);
        }
    }
};

// out = in / S[s_idx]      (gmres_mgsr.f90:325-329, :384)
struct PScale : PwBase<0> {
    const double *in;
    double *out;
    const double *S;
    int s_idx;
    FastDiv fd;
    __device__ __forceinline__ void init() { fd.set(S[s_idx]); }
    template <int VEC>
    __device__ __forceinline__ void elem(size_t i, double *) const {
        double v[VEC];
        KL_LD(VEC, v, in, i)
#pragma unroll
        for (int k = 0; k < VEC; ++k) v[k] = fd.div(v[k]);
        KL_ST(VEC, out, i, v)
    }
};

// x += alpha p ; r -= alpha ax ; acc0 = sum r*r     (cg.f90:127-133)
struct PCgUpdate : PwBase<1> {
    double *x, *r;
    const double *p, *ax;
    const double *S;
    double alpha;
    __device__ __forceinline__ void init() { alpha = S[S_ALPHA]; }
    template <int VEC>
    __device__ __forceinline__ void elem(size_t i, double *acc) const {
        double vx[VEC], vr[VEC], vp[VEC], va[VEC];
        KL_LD(VEC, vp, p, i)
        KL_LD(VEC, va, ax, i)
        if (VEC == 2) {
            double2 t = *reinterpret_cast<const double2 *>(x + i);
            vx[0] = t.x; vx[VEC - 1] = t.y;
            t = *reinterpret_cast<const double2 *>(r + i);
            vr[0] = t.x; vr[VEC - 1] = t.y;
        } else {
            vx[0] = x[i];
            vr[0] = r[i];
        }
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            vx[v] = fma(alpha, vp[v], vx[v]);
            vr[v] = fma(-alpha, va[v], vr[v]);
            acc[0] = fma(vr[v], vr[v], acc[0]);
        }
        KL_ST(VEC, x, i, vx)
        KL_ST(VEC, r, i, vr)
    }
};

// y = a
struct PCopy : PwBase<0> {
    const double *a;
    double *y;
    __device__ __forceinline__ void init() {}
    template <int VEC>
    __device__ __forceinline__ void elem(size_t i, double *) const {
        double va[VEC];
        KL_LD(VEC, va, a, i)
        KL_ST(VEC, y, i, va)
    }
};

// generic y = a + s*b with s = sign * S[s_idx]   (bicgstab.f90:132-135 s = r - alpha*ap)
struct PAxpy : PwBase<0> {
    const double *a, *b;
    double *y;
    const double *S;
    int s_idx;
    double sign, s;
    __device__ __forceinline__ void init() { s = sign * S[s_idx]; }
    template <int VEC>
    __device__ __forceinline__ void elem(size_t i, double *) const {
        double va[VEC], vb[VEC];
        KL_LD(VEC, va, a, i)
        KL_LD(VEC, vb, b, i)
#pragma unroll
        for (int v = 0; v < VEC; ++v) va[v] = fma(s, vb[v], va[v]);
        KL_ST(VEC, y, i, va)
    }
};

}  // namespace kl
