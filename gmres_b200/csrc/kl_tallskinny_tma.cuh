// kl_tallskinny_tma.cuh -- TMA-staged tall-skinny kernels for the Gram-Schmidt passes.
//
//   UPDATE = false :  h_out = V(:,0..nc-1)^T w                        (projection)
//   UPDATE = true  :  w <- w - V h_in ; h_out = V^T w (the NEW w)     (update fused with the
//                     second projection: the V tile is read from HBM once for both)
//
// A CTA streams row tiles of V (kTsRB = 32 rows x nc columns, one 2-D TMA box per tile,
// [col][row] in shared memory) through a kTsNst-deep mbarrier ring.  Thread layout:
// lane = row inside the tile, warp = column slice (kTsCpw columns).  Each thread loads its
// slice of the tile row into registers once and uses it for both the update (reduction
// across warps through shared memory) and the projection (per-thread accumulators, reduced
// across lanes only once at the end of the kernel).  Deterministic two-stage reduction as
// in k_vtw.  Requires nc <= 8*kTsCpw = 96 columns, n and ldv even.
#pragma once
#include "kl_stencil_tma.cuh"

namespace kl {

constexpr int kTsRB = 32;    // rows per warp-row-group; a tile has kTsRB * RM rows (RM = 1,2,4,8)
constexpr int kTsNst = 4;    // ring depth
// row-group multiplier: few columns -> taller tiles, so that a tile stays ~16-24 KB and the
// eight warps split into RM row groups x 8/RM column slices
inline int ts_rm(int nc) { return nc <= 12 ? 8 : (nc <= 24 ? 4 : (nc <= 48 ? 2 : 1)); }

// Optional epilogue of the projection kernel's last block (compact-WY Householder GMRES, kl_hh.cu):
//   tvec(0..nc-1) = T(0..nc-1, 0..nc-1)^T out(0..nc-1)   -- the O(j^2) triangular product that sits between the two
// tall-skinny passes of a step; run by the eight warps of the last block instead of a kernel of its own.
struct TsTail {
    const double *T;    // nullptr: no epilogue
    double *tvec;
    int ldt;
};
static __device__ __noinline__ void ts_tail_tT(const TsTail tt, const double *s, const int nc) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int c = wid; c < nc; c += kTsWarps) {
        double t = 0.0;
        for (int r = lane; r <= c; r += 32) t = fma(tt.T[(size_t)c * tt.ldt + r], s[r], t);
        t = warp_sum(t);
        if (lane == 0) tt.tvec[c] = t;
    }
}

constexpr int kTsLd = 128;                 // doubles per row of partial sums (m + 2 <= 128)
constexpr int kTsGroup = 16;               // CTAs per first-level reduction group (296 CTAs -> 19 groups: one round each level)
constexpr int kTsGroupRow = 512;           // rows kTsGroupRow.. of the partials hold the group sums (grid <= 512 CTAs)
constexpr int kTsGroupCounterOffset = 15;  // group arrival counters: counter[15 .. 15 + 64) (counter = Ctx::d_counter + 1)

#ifdef KL_TRACE
// debug build only: per-CTA time stamps of the last tall-skinny pass of this translation unit
//   [0] entry  [1] first tile landed  [2] tile loop done  [3] arrived at the grid counter  [4] (last block) sums written
static __device__ unsigned long long g_trace_ts[8 * 512];
#define KL_TS_STAMP(k) if (threadIdx.x == 0 && blockIdx.x < 512) g_trace_ts[8 * blockIdx.x + (k)] = gtimer();
#else
#define KL_TS_STAMP(k)
#endif

// one tall-skinny pass by the whole grid (body of k_ts_tma; also called three times in a row, with grid-wide
// barriers in between, by the cooperative kernel k_cgs2_coop in kl_gmres.cu)
template <bool UPDATE>
__device__ __forceinline__ void ts_pass(const CUtensorMap &tmV, double *w, const size_t n, const int nc, const int RM,
                                        const double *__restrict__ h_in, double *__restrict__ partials,
                                        unsigned int *counter, double *__restrict__ out, const GmresDev &G, const int j,
                                        const int h_mode, const long long tail0, const TsTail &tt,
                                        unsigned char *smem_raw, unsigned int *sync_flag = nullptr,
                                        const unsigned sync_target = 0u) {
    KL_TS_STAMP(0)
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int RB = kTsRB * RM;                                  // rows per tile
    const int nslice = kTsWarps / RM;                            // column slices
    const int rowgrp = wid % RM, slice = wid / RM;
    const unsigned tile_doubles = (unsigned)RB * nc;
    const unsigned tile_bytes = tile_doubles * sizeof(double);
    const unsigned tile_stride = (tile_bytes + 127u) & ~127u;
    double *tiles = reinterpret_cast<double *>(smem_raw);
    unsigned char *p = smem_raw + (size_t)kTsNst * tile_stride;
    unsigned long long *full = reinterpret_cast<unsigned long long *>(p);
    p += 64;
    double *s_h = reinterpret_cast<double *>(p);               // nc (padded to 96)
    p += 96 * sizeof(double);
    double *s_part = reinterpret_cast<double *>(p);            // [slice][RB rows] = 256 doubles
    p += 8 * kTsRB * sizeof(double);
    double *s_red = reinterpret_cast<double *>(p);             // [rowgrp * 8 + slice][kTsCpw]
    __shared__ int s_last;

    const size_t ntiles = (n + RB - 1) / RB;
    const int c0 = slice * kTsCpw;                             // first column of this warp's slice
    int ncw = nc - c0;                                         // columns in this slice
    ncw = ncw < 0 ? 0 : (ncw > kTsCpw ? kTsCpw : ncw);

    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < kTsNst; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (UPDATE)
        for (int c = threadIdx.x; c < nc; c += kTsThreads) s_h[c] = h_in[c];
    __syncthreads();
    auto issue = [&](size_t t, int s) {
        mbar_expect_tx(&full[s], tile_bytes);
        tma_load_2d(reinterpret_cast<unsigned char *>(tiles) + (size_t)s * tile_stride, &tmV, &full[s],
                    (int)(t * RB), 0);
    };
    // tiles of this CTA: t = blockIdx.x + k * gridDim.x
    size_t my_tiles = ntiles > blockIdx.x ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    if (threadIdx.x == 0) {
        for (size_t k = 0; k < (size_t)kTsNst && k < my_tiles; ++k) issue(blockIdx.x + k * gridDim.x, (int)k);
    }
    double acc[kTsCpw];
#pragma unroll
    for (int c = 0; c < kTsCpw; ++c) acc[c] = 0.0;
    double nacc = 0.0;
    // w of the first tile (register prefetch, one tile ahead)
    const int rin = rowgrp * kTsRB + lane;                     // row inside the tile
    size_t row = (size_t)blockIdx.x * RB + rin;
    double wnext = (my_tiles > 0 && row < n) ? w[row] : 0.0;

    for (size_t k = 0; k < my_tiles; ++k) {
        const int s = (int)(k % kTsNst);
        const size_t t = blockIdx.x + k * gridDim.x;
        row = t * RB + rin;
        double wr = wnext;
        {
            const size_t rn = (t + gridDim.x) * RB + rin;
            wnext = (k + 1 < my_tiles && rn < n) ? w[rn] : 0.0;
        }
        mbar_wait(&full[s], (unsigned)((k / kTsNst) & 1));
#ifdef KL_TRACE
        if (k == 0) { KL_TS_STAMP(1) }
#endif
        const double *tile = reinterpret_cast<const double *>(reinterpret_cast<unsigned char *>(tiles) +
                                                               (size_t)s * tile_stride);
        double v[kTsCpw];
#pragma unroll
        for (int c = 0; c < kTsCpw; ++c) v[c] = (c < ncw) ? tile[(size_t)(c0 + c) * RB + rin] : 0.0;
        if (UPDATE) {
            double part = 0.0;
#pragma unroll
            for (int c = 0; c < kTsCpw; ++c)
                if (c < ncw) part = fma(s_h[c0 + c], v[c], part);
            s_part[slice * RB + rin] = part;
            __syncthreads();
            double sum = 0.0;
            for (int q = 0; q < nslice; ++q) sum += s_part[q * RB + rin];
            wr = wr - sum;
            if (slice == 0 && row < n) w[row] = wr;
        }
#pragma unroll
        for (int c = 0; c < kTsCpw; ++c) acc[c] = fma(v[c], wr, acc[c]);
        if (tail0 >= 0 && slice == 0 && (long long)row >= tail0 && row < n) nacc = fma(wr, wr, nacc);
        __syncthreads();   // tile s (and s_part) consumed by every thread
        if (threadIdx.x == 0 && k + kTsNst < my_tiles) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            issue(blockIdx.x + (k + kTsNst) * gridDim.x, s);
        }
    }
    KL_TS_STAMP(2)
    // ---- block stage: reduce over the 32 row-lanes, then over the RM row groups
#pragma unroll
    for (int c = 0; c < kTsCpw; ++c) {
        double sv = warp_sum(acc[c]);
        if (lane == 0) s_red[(rowgrp * 8 + slice) * kTsCpw + c] = sv;
    }
    __syncthreads();
    // partial sums: one row of kTsLd doubles per CTA (column-contiguous: the writes below and the one-thread-per-column
    // sums of the grid stage are coalesced)
    if (threadIdx.x < nc) {
        const int sl = threadIdx.x / kTsCpw, cc = threadIdx.x % kTsCpw;
        double sv = 0.0;
        for (int q = 0; q < RM; ++q) sv += s_red[(q * 8 + sl) * kTsCpw + cc];
        partials[(size_t)blockIdx.x * kTsLd + threadIdx.x] = sv;
    }
    if (tail0 >= 0) {
        __syncthreads();
        double sv = warp_sum(nacc);
        if (lane == 0) s_red[wid] = sv;     // only slice-0 warps hold non-zero values
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int q = 0; q < kTsWarps; ++q) t += s_red[q];
            partials[(size_t)blockIdx.x * kTsLd + nc] = t;
        }
    }
    // ---- grid stage, two levels, fixed order (same launch configuration => same bits).  Level 1: the CTAs form groups
    // of kTsGroup consecutive block ids; the last CTA of a group to arrive adds the group's partials (one thread per
    // column, all loads in flight) -- this runs while other groups are still streaming.  Level 2: the last group to
    // finish adds the <= 19 group sums the same way.  The flat version (one warp per column pair, lanes over the 296
    // CTAs) took 6 us at m = 95: six serial rounds of L2 latency, a third of the whole pass on L2-resident bases.
    const int ncr = nc + (tail0 >= 0 ? 1 : 0);
    const unsigned grp = blockIdx.x / kTsGroup, ngrp = (gridDim.x + kTsGroup - 1) / kTsGroup;
    const unsigned gsize = min((unsigned)kTsGroup, gridDim.x - grp * kTsGroup);
    unsigned int *gcounter = counter + kTsGroupCounterOffset;
    auto wait_result = [&]() {
        // persistent use (k_cgs2_coop): wait until the last block has published the column sums.  The arrivals and
        // this flag together are the grid barrier -- no second barrier after the reduction.
        if (sync_flag) {
            if (threadIdx.x == 0) {
                unsigned v;
                do {
                    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(sync_flag) : "memory");
                } while ((int)(v - sync_target) < 0);
            }
            __syncthreads();
        }
    };
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        unsigned prev = atomicAdd(gcounter + grp, 1u);
        s_last = (prev == gsize - 1);
        KL_TS_STAMP(3)
    }
    __syncthreads();
    if (!s_last) {
        wait_result();
        return;
    }
    __threadfence();
    if ((int)threadIdx.x < ncr) {
        const double *pp = partials + (size_t)grp * kTsGroup * kTsLd + threadIdx.x;
        double t[kTsGroup];
#pragma unroll
        for (int q = 0; q < kTsGroup; ++q) t[q] = (unsigned)q < gsize ? __ldcg(pp + (size_t)q * kTsLd) : 0.0;
        double sv = 0.0;
#pragma unroll
        for (int q = 0; q < kTsGroup; ++q) sv += t[q];
        partials[(size_t)(kTsGroupRow + grp) * kTsLd + threadIdx.x] = sv;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        gcounter[grp] = 0u;
        __threadfence();
        unsigned prev = atomicAdd(counter, 1u);
        s_last = (prev == ngrp - 1);
    }
    __syncthreads();
    if (!s_last) {
        wait_result();
        return;
    }
    __threadfence();
    if ((int)threadIdx.x < ncr) {
        const double *pp = partials + (size_t)kTsGroupRow * kTsLd + threadIdx.x;
        double sv = 0.0;
        for (unsigned g0 = 0; g0 < ngrp; g0 += 20) {        // <= 19 groups at 296 CTAs: one round of loads in flight
            double t[20];
#pragma unroll
            for (int q = 0; q < 20; ++q) t[q] = g0 + q < ngrp ? __ldcg(pp + (size_t)(g0 + q) * kTsLd) : 0.0;
#pragma unroll
            for (int q = 0; q < 20; ++q) sv += t[q];
        }
        const int col = threadIdx.x;
        out[col] = sv;
        if (h_mode && col < nc) {
            double *Hj = G.H + (size_t)j * G.ldh;
            Hj[col] = (h_mode == 2) ? Hj[col] + sv : sv;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) *counter = 0u;
    KL_TS_STAMP(4)
    if (!UPDATE && tt.T) {
        __syncthreads();          // out[] was written by this block's warps
        ts_tail_tT(tt, out, nc);
    }
    if (sync_flag) {
        __syncthreads();          // every warp's column sums (and the counter reset) are written
        if (threadIdx.x == 0) {
            __threadfence();
            asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(sync_flag), "r"(sync_target) : "memory");
        }
    }
}

template <bool UPDATE>
__global__ void __launch_bounds__(kTsThreads, 2)
k_ts_tma(const __grid_constant__ CUtensorMap tmV, double *w, const size_t n, const int nc, const int RM,
         const double *__restrict__ h_in, double *__restrict__ partials, unsigned int *counter,
         double *__restrict__ out, const GmresDev G, const int j, const int h_mode,
         const int *__restrict__ flags, const long long tail0 /* >= 0: out[nc] = sum_{row >= tail0} w_row^2 */,
         const TsTail tt) {
    griddep_wait();      // before the gate: flags, w, h_in and the newest column of V come from the predecessors
    griddep_launch();
    if (flags && flags[I_CONV_AT] >= 0) return;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    ts_pass<UPDATE>(tmV, w, n, nc, RM, h_in, partials, counter, out, G, j, h_mode, tail0, tt, smem_raw);
}

inline size_t ts_tma_smem(int nc) {
    const size_t tile_stride = ((size_t)kTsRB * ts_rm(nc) * nc * sizeof(double) + 127) & ~size_t(127);
    return kTsNst * tile_stride + 64 + 96 * sizeof(double) + 8 * kTsRB * sizeof(double) +
           64 * kTsCpw * sizeof(double) + 128;
}

// host: 2-D FP64 tensor map over V (n rows fastest, ncols columns, ld = ldv), box kTsRB x nc
int tmap_encode_v(Ctx *c, CUtensorMap *out, const double *V, size_t n, size_t ldv, int ncols_total, int nc);

}  // namespace kl
