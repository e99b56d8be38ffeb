// kl_stencil_tma.cuh -- TMA-staged variant of the marching stencil kernel (sm_100a).
//
// Same functor contract and arithmetic as k_stencil (kl_internal.cuh); the only
// difference is where the input lines come from.  Each CTA owns a strip of
// kTmaStrip = 252 columns and marches down `rows` grid lines.  One elected thread
// feeds a ring of kTmaStages shared-memory stages with 2-D TMA tile loads
// (cp.async.bulk.tensor.2d, box = 256 columns x kTmaSR lines per input array,
// completion on an mbarrier), so the bytes in flight live in shared memory
// instead of registers: ~3 stages x 4-12 KB per CTA, 4-8 CTAs per SM.  The box
// starts two columns left of the strip (16-byte aligned LDS.128 for every
// thread) and one line above it; TMA's out-of-bounds zero fill supplies the
// zero-Dirichlet boundary on all four sides with no branches.  Thread t owns box
// columns (2t, 2t+1); threads 0 and 127 only carry halo columns.  Left/right
// neighbours come from warp shuffles (warp-edge lanes read one extra scalar from
// shared memory), up/down from registers, exactly as in k_stencil.
#pragma once
#include <cuda.h>

namespace kl {

constexpr int kTmaBoxX = 256;
constexpr int kTmaStrip = 252;
constexpr int kTmaSR = 2;       // grid lines per stage
constexpr int kTmaStages = 4;

template <int NIN>
struct alignas(64) TMaps {
    CUtensorMap m[NIN];
};

__device__ __forceinline__ unsigned smem_u32(const void *p) {
    return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "KL_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra KL_DONE;\n"
        "bra KL_WAIT;\n"
        "KL_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *tm, unsigned long long *bar,
                                            int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<unsigned long long>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}

#ifdef KL_TRACE
// debug build only: per-CTA (start, end, SM id) time stamps of the last k_stencil_tma launch of this translation unit
static __device__ unsigned long long g_trace[3 * 16384 + 8];
__device__ __forceinline__ unsigned long long gtimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ unsigned smid() {
    unsigned r;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(r));
    return r;
}
#endif

template <class F, int OPK>
__global__ void __launch_bounds__(kStencilThreads)
k_stencil_tma(const F f_in, const Geo g, const RedCtl rc, const PostAny post, const int fuse_post,
              const __grid_constant__ TMaps<F::NIN> tm) {
    // programmatic dependent launch (see griddep_wait): kLateWait functors run their whole prologue, the first TMA
    // stages and the first lines of u before they need anything the preceding kernel produces
#ifdef KL_TRACE
    const unsigned long long t_start = gtimer();
#endif
    if (!F::kLateWait) griddep_wait();
    griddep_launch();
#ifdef KL_TRACE
    const unsigned long long t_go = gtimer();
#endif
    if (f_in.skip()) {
        if (F::kLateWait) griddep_wait();   // a grid completes only after its predecessor did
        return;
    }
    F f = f_in;
    f.init();
    if (!F::kLateWait) f.late_init();
    bool waited = !F::kLateWait;
    constexpr int NIN = F::NIN;
    constexpr int NRED = F::NRED;
    constexpr int NR = NRED > 0 ? NRED : 1;
    constexpr int VEC = 2;
    constexpr int SR = kTmaSR, NST = kTmaStages;
    constexpr unsigned kStageDoubles = NIN * SR * kTmaBoxX;
    constexpr unsigned kStageBytes = kStageDoubles * sizeof(double);

    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *sbuf = reinterpret_cast<double *>(smem_raw);
    unsigned long long *full = reinterpret_cast<unsigned long long *>(smem_raw + (size_t)NST * kStageBytes);

    // PERSISTENT CTAs: the grid holds at most as many CTAs as are resident at once (launch_stencil); CTA c processes
    // the tiles c, c + gridDim.x, c + 2 gridDim.x, ... (tile = strip + gx * line band, strips fastest).  The TMA ring
    // runs ACROSS tile boundaries: while the last stages of one tile are consumed, the first stages of the next one
    // are already in flight.  With one CTA per tile the 4-5 "waves" of a slab of the 8-GPU strong-scaling run each
    // started and drained in lockstep, and every wave boundary was a few microseconds of idle memory system
    // (profiles/r02_cta_timeline_k1.txt).
    const int tid = threadIdx.x, lane = tid & 31;
    const bool has_l = lane == 0 && tid > 0;
    const bool has_r = lane == 31 && tid < kStencilThreads - 1;
    const int ntiles = g.gx * g.gy;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < NST; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // producer state (used by thread 0 only): the next stage to issue is stage p_k of tile p_tile
    int p_tile = blockIdx.x, p_k = 0, p_i0 = 0, p_jstart = 0, p_nst = 0;
    unsigned p_cnt = 0;
    auto p_geometry = [&]() {
        if (p_tile < ntiles) {
            int a0, a1;
            tile_lines(g, p_tile / g.gx, a0, a1);
            p_i0 = (p_tile % g.gx) * kTmaStrip;
            p_jstart = a0 - 1;
            p_nst = (a1 - a0 + 2 + SR - 1) / SR;
        }
    };
    auto issue_next = [&]() {
        if (p_tile >= ntiles) return;
        const int s = (int)(p_cnt % NST);
        mbar_expect_tx(&full[s], kStageBytes);
#pragma unroll
        for (int a = 0; a < NIN; ++a)
            tma_load_2d(sbuf + (size_t)s * kStageDoubles + (size_t)a * SR * kTmaBoxX, &tm.m[a], &full[s],
                        p_i0 - 2, p_jstart + p_k * SR);
        ++p_cnt;
        if (++p_k == p_nst) {
            p_tile += gridDim.x;
            p_k = 0;
            p_geometry();
        }
    };
    if (tid == 0) {
        p_geometry();
        for (int k = 0; k < NST; ++k) issue_next();
    }

    double acc[NR];
#pragma unroll
    for (int k = 0; k < NR; ++k) acc[k] = 0.0;
    unsigned c_cnt = 0;      // stages consumed so far (slot = c_cnt % NST, parity = (c_cnt / NST) & 1)

    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int i0 = (tile % g.gx) * kTmaStrip;    // first output column of the strip
    const int gi = i0 - 2 + 2 * tid;             // global column of this thread's pair
    const bool outp = tid >= 1 && tid <= kStencilThreads - 2 && gi < g.nx;
    int j0, j1;
    tile_lines(g, tile / g.gx, j0, j1);
    const int jstart = j0 - 1;
    const int nrows = j1 - j0 + 2;
    const int nst = (nrows + SR - 1) / SR;
    double up[VEC] = {0.0, 0.0}, cu[VEC] = {0.0, 0.0}, dn[VEC], cl = 0.0, cr = 0.0, dl, dr;
    double rawCu[NIN][VEC];
#pragma unroll
    for (int a = 0; a < NIN; ++a) rawCu[a][0] = rawCu[a][1] = 0.0;

    const bool patch_lo = f.lo[0] != nullptr && j0 == 0;
    const bool patch_hi = f.hi[0] != nullptr && j1 == g.ny;
    const int k_hi = (g.ny - jstart) / SR, rr_hi = (g.ny - jstart) % SR;

    for (int k = 0; k < nst; ++k, ++c_cnt) {
        const int s = (int)(c_cnt % NST);
        mbar_wait(&full[s], (unsigned)((c_cnt / NST) & 1));
        double *st = sbuf + (size_t)s * kStageDoubles;
        // multi-GPU: the lines above / below this rank's slab come from the neighbours' halo buffers
        if ((patch_lo && k == 0) || (patch_hi && k == k_hi)) {
            const bool lo = patch_lo && k == 0;
            const int rr = lo ? 0 : rr_hi;
#pragma unroll
            for (int a = 0; a < NIN; ++a) {
                const double *src = lo ? f.lo[a] : f.hi[a];
#pragma unroll
                for (int e = 0; e < VEC; ++e) {
                    const int c = gi + e;
                    st[((size_t)a * SR + rr) * kTmaBoxX + 2 * tid + e] = (c >= 0 && c < g.nx) ? __ldg(src + c) : 0.0;
                }
            }
            __syncthreads();
        }
#pragma unroll
        for (int rr = 0; rr < SR; ++rr) {
            const int r = jstart + k * SR + rr;
            if (r <= j1) {
                double raw[NIN][VEC], rl[NIN], rrg[NIN];
#pragma unroll
                for (int a = 0; a < NIN; ++a) {
                    const double *row = st + ((size_t)a * SR + rr) * kTmaBoxX;
                    const double2 t = *reinterpret_cast<const double2 *>(row + 2 * tid);
                    raw[a][0] = t.x;
                    raw[a][1] = t.y;
                    rl[a] = has_l ? row[2 * tid - 1] : 0.0;
                    rrg[a] = has_r ? row[2 * tid + 2] : 0.0;
                }
#pragma unroll
                for (int v = 0; v < VEC; ++v) {
                    double t[NIN];
#pragma unroll
                    for (int a = 0; a < NIN; ++a) t[a] = raw[a][v];
                    dn[v] = f.point(t);
                }
                dl = has_l ? f.point(rl) : 0.0;
                dr = has_r ? f.point(rrg) : 0.0;
                if (r >= j0 + 1) {
                    double l = __shfl_up_sync(0xffffffffu, cu[1], 1);
                    double rt = __shfl_down_sync(0xffffffffu, cu[0], 1);
                    if (lane == 0) l = cl;
                    if (lane == 31) rt = cr;
                    if (F::kLateWait && !waited) {
                        griddep_wait();
                        f.late_init();
                        waited = true;
                    }
                    if (outp) {
                        double au[VEC];
                        au[0] = apply5<OPK>(cu[0], l, cu[1], dn[0], up[0], f.coef);
                        au[1] = apply5<OPK>(cu[1], cu[0], rt, dn[1], up[1], f.coef);
                        if constexpr (F::kPush)
                            f.template store<VEC>((size_t)(r - 1) * g.nx + gi, rawCu, cu, au, acc,
                                                  (r == 1 ? 1 : 0) | (r == g.ny ? 2 : 0));
                        else
                            f.template store<VEC>((size_t)(r - 1) * g.nx + gi, rawCu, cu, au, acc);
                    }
                }
#pragma unroll
                for (int v = 0; v < VEC; ++v) {
                    up[v] = cu[v];
                    cu[v] = dn[v];
                }
                cl = dl;
                cr = dr;
#pragma unroll
                for (int a = 0; a < NIN; ++a) {
                    rawCu[a][0] = raw[a][0];
                    rawCu[a][1] = raw[a][1];
                }
            }
        }
        __syncthreads();   // every thread has read stage s
        if (tid == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            issue_next();  // refills slot s with the stage NST ahead (of this tile or of the CTA's next one)
        }
    }
    }   // tiles of this CTA
#ifdef KL_TRACE
    if (threadIdx.x == 0) {
        const unsigned b = blockIdx.x;
        if (b < 16384) {
            g_trace[3 * b] = t_start;
            g_trace[3 * b + 1] = gtimer();
            g_trace[3 * b + 2] = ((unsigned long long)smid() << 48) | (t_go - t_start);
        }
    }
#endif
    if (NRED > 0) {
        __shared__ double sm[NR * (kStencilThreads / 32)];
        __shared__ int s_flag;
        block_sum<NR, kStencilThreads>(acc, sm);
        const unsigned nb = gridDim.x, bid = blockIdx.x;
        if (F::kLateWait && !waited) griddep_wait();
        if (grid_sum<NR, kStencilThreads>(acc, rc, nb, bid, &s_flag, sm)) {
            if (rc.peer) peer_allreduce_block<NR>(rc);
            if (fuse_post && threadIdx.x == 0) post.run();
#ifdef KL_TRACE
            if (threadIdx.x == 0) g_trace[3 * 16384] = gtimer();     // end of the last block's tail
#endif
        }
    }
}

// host: encode a 2-D FP64 tensor map (nx x ny, box 256 x kTmaSR); implemented in kl_core.cu
int tmap_encode(Ctx *c, CUtensorMap *out, const double *base, int nx, int ny);

template <int NIN>
constexpr size_t tma_smem_bytes() {
    return (size_t)kTmaStages * NIN * kTmaSR * kTmaBoxX * sizeof(double) + 64;
}

}  // namespace kl
