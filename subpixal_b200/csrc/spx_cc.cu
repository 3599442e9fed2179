// Fused gather + 4-phase cross-correlation + interlace + sub-pixel fit kernel
// and its C-ABI entry points (2) spx_cc_fit_stack and (3) spx_cc_fit_frames.
//
// One persistent CTA processes one cutout pair at a time: the five tiles are
// read from HBM exactly once per pass over them, every intermediate (spectra,
// cross-power, correlation rows) lives in shared memory (or, for cutouts too
// large for one SM, in a per-CTA slab that stays L2 resident), and only the
// (dx, dy) pair goes back to HBM unless the caller asks for the surface.
#include <cooperative_groups.h>
#include <cuda.h>               // CUtensorMap (types only; the encoder is fetched at run time)
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include "../../include/subpixal_b200.h"
#include "spx_cc_engine.cuh"

namespace cg = cooperative_groups;

namespace spx {

struct CcParams {
    const float *src[5];        // stack bases or frame bases
    long long pair_stride[5];   // elements between pairs (0 in frame mode)
    long long pitch;            // row pitch of every source (elements)
    int H, W;                   // source image size
    const int *boxes;           // [n,4] or nullptr (stack mode)
    // ragged mode (shapes != nullptr): pair p = order[k] has shape shapes[p], its five
    // tiles [5][ny][nx] start at src[0] + offsets[p]; ny/nx below are the largest shape
    const int *shapes; const long long *offsets; const int *order; const long long *icc_offsets;
    int n, ny, nx, cc_type;
    unsigned flags;
    double *shifts; int *peak; int *status; float *icc_out;
    unsigned char *workspace; size_t slab_bytes;
    size_t off_icc, off_w, off_r;   // byte offsets inside a slab
    // staged configurations, frame mode: one 2-D tensor map per frame (box NMAX x NMAX,
    // out-of-bounds elements read as zero = Cutout 'fill' mode), in global memory
    const CUtensorMap *tmap;
    // dynamic work distribution (single-CTA classes): pairs beyond the first two of a CTA
    // are claimed from this counter (zeroed before the launch), so that a CTA that starts
    // late - e.g. because an NCCL kernel still occupies its SM - simply takes fewer pairs
    unsigned *claim;
};

// ---- configurations (one per supported maximum cutout size) -----------------
//                P   E1  E2 NMAX   T  MINB  W_SMEM R_SMEM INPLACE
// tuning knobs (tools/variant.sh builds variants with -D overrides)
#ifndef SPX32_E1
#define SPX32_E1 6
#endif
#ifndef SPX32_T
#define SPX32_T 128
#endif
#ifndef SPX32_MINB
#define SPX32_MINB 6
#endif
#ifndef SPX64_E1
#define SPX64_E1 12
#endif
#ifndef SPX64_T
#define SPX64_T 256
#endif
#ifndef SPX64_MINB
#define SPX64_MINB 2
#endif
// Intermediate classes with a power-of-two transform (P >= 1.5 n): pure radix-2/4 codelets,
// generic (runtime-geometry) path only.  They exist for ragged batches - the bounding-box
// cutouts of create_cutouts (cutout.py:159-175) are mostly 15-50 px wide.
#ifndef SPX20_T
#define SPX20_T 64
#endif
#ifndef SPX20_MINB
#define SPX20_MINB 12
#endif
#ifndef SPX42_T
#define SPX42_T 128
#endif
#ifndef SPX42_MINB
#define SPX42_MINB 4
#endif
#ifndef SPX84_T
#define SPX84_T 512
#endif
using Cfg20  = CcCfg< 32,  4,  8,  20, SPX20_T, SPX20_MINB, true,  true,  true>;
using Cfg42  = CcCfg< 64,  8,  8,  42, SPX42_T, SPX42_MINB, true,  true,  true>;
using Cfg84  = CcCfg<128,  8, 16,  84, SPX84_T, 1, true,  true,  true>;
// 170 / 256 px classes: W and R either in a per-CTA global slab (default) or in the shared
// memory of a 4- / 8-CTA cluster (row split of W and R like the 128 px class, mirrored R
// rows over DSMEM).  The cluster forms move exactly the algorithmic bytes through DRAM
// (1.0x instead of 14.5x at 256 px) but spend a third of their time in the 9 cluster-wide
// barriers per pair (the slowest of 8 CTAs sets the pace of every pass): 139 k vs 177 k
// pairs/s at 256 px, 354 k vs 391 k at 170 px (profiles/r2_cc_fit_cfg256_cluster8_*), so
// the slab forms stay the default; -DSPX256_CLUSTER=1 / -DSPX170_CLUSTER=1 select the
// cluster forms (parity-tested on the GPU and by tests/test_emulator.py).
#ifndef SPX170_CLUSTER
#define SPX170_CLUSTER 0
#endif
#ifndef SPX170_T
#define SPX170_T (SPX170_CLUSTER ? 512 : 256)
#endif
#if SPX170_CLUSTER
using Cfg170 = CcCfg<256, 16, 16, 170, SPX170_T, 1, true, true, true, 4>;  // 4-CTA cluster, DSMEM
#else
#ifndef SPX170_MINB
#define SPX170_MINB 2
#endif
using Cfg170 = CcCfg<256, 16, 16, 170, SPX170_T, SPX170_MINB, false, false>;   // W / R in a global slab
#endif
using Cfg32  = CcCfg< 48, SPX32_E1, 48 / SPX32_E1,  32, SPX32_T, SPX32_MINB, true,  true,  true>;
using Cfg64  = CcCfg< 96, SPX64_E1, 96 / SPX64_E1,  64, SPX64_T, SPX64_MINB, true,  true,  true>;
// TMA-staged forms of the two headline classes (tiles of the next pair land in shared memory
// while the current pair is transformed): 20 KB / 80 KB more shared memory per CTA, hence
// 4 instead of 6 CTAs per SM at 32 px and ONE CTA of 16 warps instead of two of 8 at 64 px.
// Selected at run time by SPX_STAGE=1 for uniform full-size batches (see use_stage()).
#ifndef SPX32S_MINB
#define SPX32S_MINB 4
#endif
#ifndef SPX64S_T
#define SPX64S_T 512
#endif
using Cfg32S = CcCfg< 48, SPX32_E1, 48 / SPX32_E1,  32, SPX32_T, SPX32S_MINB, true, true, true, 1, true>;
using Cfg64S = CcCfg< 96, SPX64_E1, 96 / SPX64_E1,  64, SPX64S_T, 1, true, true, true, 1, true>;
#ifndef SPX128_CLUSTER
#define SPX128_CLUSTER 1
#endif
#if SPX128_CLUSTER
#ifndef SPX128_T
#define SPX128_T 512
#endif
#ifndef SPX128_INPLACE
#define SPX128_INPLACE 1
#endif
using Cfg128 = CcCfg<192, 12, 16, 128, SPX128_T, 1, true,  true,  SPX128_INPLACE != 0, 2>;   // 2-CTA cluster, DSMEM
#else
// (build knob: the 128 px class in the fused global-slab form of the 170 / 256 px classes)
#ifndef SPX128_SLAB_T
#define SPX128_SLAB_T 256
#endif
#ifndef SPX128_SLAB_MINB
#define SPX128_SLAB_MINB 2
#endif
using Cfg128 = CcCfg<192, 12, 16, 128, SPX128_SLAB_T, SPX128_SLAB_MINB, false, false>;
#endif
#ifndef SPX256_CLUSTER
#define SPX256_CLUSTER 0
#endif
#if SPX256_CLUSTER
#ifndef SPX256_T
#define SPX256_T 384
#endif
using Cfg256 = CcCfg<384, 24, 16, 256, SPX256_T, 1, true, true, true, 8>;  // 8-CTA cluster, DSMEM
#else
// (two CTAs of 8 warps per SM: with the fused column passes the only shared memory is the
// 98.6 KB scratch S, and two CTAs hide each other's copy passes - 210 k vs 191 k pairs/s for
// one CTA of 16 warps, 180 k for the unfused form)
#ifndef SPX256_T
#define SPX256_T 256
#endif
#ifndef SPX256_MINB
#define SPX256_MINB 2
#endif
using Cfg256 = CcCfg<384, 24, 16, 256, SPX256_T, SPX256_MINB, false, false>;   // W / R in a global slab
#endif

template <class C> struct Layout {
    // shared memory carve-up (bytes); everything 16-byte aligned
    static constexpr size_t off_ps = 0;
    static constexpr size_t off_red = 256;                       // reductions
    // [NW][8] doubles (statistics); the argmax reduction keeps its indices at byte 256
    static constexpr size_t red_bytes = C::NW * 64 > 512 ? C::NW * 64 : 512;
    static constexpr size_t off_xchg = off_red + red_bytes;       // Xchg (whole only in cluster mode)
    static constexpr size_t xchg_bytes = C::CL > 1 ? 656 : 64;
    static constexpr size_t off_win = off_xchg + xchg_bytes;      // FitWin (surface kept in W)
    static constexpr size_t win_bytes = SurfaceInW<C>::value ? 832 : 0;
    static constexpr size_t off_rrow = off_win + win_bytes;
    static constexpr size_t rrow_bytes = ((C::RH * 2 + 15) / 16) * 16;
    static constexpr size_t off_xrow = off_rrow + rrow_bytes;
    static constexpr size_t xrow_bytes = 0;
    static constexpr size_t off_S = off_xrow + xrow_bytes;
    static constexpr size_t off_W = off_S + C::S_ELEMS * 8;
    static constexpr size_t off_R = off_W + (C::W_SMEM ? C::W_ELEMS * 8 : 0);
    // staged tiles [5][NMAX][NMAX] float (128-byte aligned for the TMA) + their mbarrier
    static constexpr size_t off_tiles = (off_R + (C::R_SMEM ? C::R_ELEMS * 8 : 0) + 127) / 128 * 128;
    static constexpr size_t off_mbar = off_tiles + C::TILE_BYTES;
    static constexpr size_t total = C::STAGE ? off_mbar + 16 : off_R + (C::R_SMEM ? C::R_ELEMS * 8 : 0);
    static_assert(sizeof(PairState) <= 256, "PairState must fit its slot");
    static_assert(sizeof(Xchg) <= 656 && offsetof(Xchg, stat) <= 64, "Xchg must fit its slot");
    static_assert(sizeof(FitWin) <= 832, "FitWin must fit its slot");
};

__device__ __forceinline__ Best warp_best(Best b) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        Best t;
        t.v = __shfl_xor_sync(0xffffffffu, b.v, o);
        t.idx = __shfl_xor_sync(0xffffffffu, b.idx, o);
        if (best_better(t, b)) b = t;
    }
    return b;
}

// Execution model of process_pair() on the device: one CTA, block barriers.
template <class C>
struct DevExec {
    double *red;     // [32][8] doubles of reduction scratch in shared memory

    template <class F> __device__ __forceinline__ void par(F f) {
        f(int(threadIdx.x));
        __syncthreads();
    }
    template <class F> __device__ __forceinline__ void single(F f) {
        if (threadIdx.x == 0) f();
        __syncthreads();
    }
    template <class F> __device__ __forceinline__ StatPartial reduce_stats(F f) {
        StatPartial st = f(int(threadIdx.x));
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            st.cnt += __shfl_xor_sync(0xffffffffu, st.cnt, o);
            st.sum += __shfl_xor_sync(0xffffffffu, st.sum, o);
            st.sq += __shfl_xor_sync(0xffffffffu, st.sq, o);
            st.rcnt += __shfl_xor_sync(0xffffffffu, st.rcnt, o);
            st.rsum += __shfl_xor_sync(0xffffffffu, st.rsum, o);
            st.rsq += __shfl_xor_sync(0xffffffffu, st.rsq, o);
            st.flags |= __shfl_xor_sync(0xffffffffu, st.flags, o);
        }
        if (lane == 0) {
            double *r = red + warp * 8;
            r[0] = st.cnt; r[1] = st.sum; r[2] = st.sq;
            r[3] = st.rcnt; r[4] = st.rsum; r[5] = st.rsq;
            r[6] = double(st.flags);
        }
        __syncthreads();
        StatPartial tot = stat_zero();
        if (threadIdx.x == 0) {
            for (int w = 0; w < C::NW; ++w) {
                const double *r = red + w * 8;
                tot.cnt += r[0]; tot.sum += r[1]; tot.sq += r[2];
                tot.rcnt += r[3]; tot.rsum += r[4]; tot.rsq += r[5];
                tot.flags |= int(r[6]);
            }
        }
        return tot;
    }
    template <class F> __device__ __forceinline__ void parts(int n, F f) {
        if (int(threadIdx.x) < n) f(int(threadIdx.x));
        __syncthreads();
    }
    template <class F, class A>
    __device__ __forceinline__ void reduce_best_into(Best *dst, F f, A after) {
        Best b = warp_best(f(int(threadIdx.x)));
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        float *redf = reinterpret_cast<float *>(red);
        int *redi = reinterpret_cast<int *>(red) + 64;
        if (lane == 0) { redf[warp] = b.v; redi[warp] = b.idx; }
        __syncthreads();
        if (threadIdx.x < 32) {
            Best t = best_init();
            if (lane < C::NW) { t.v = redf[lane]; t.idx = redi[lane]; }
            t = warp_best(t);
            if (lane == 0 && best_better(t, *dst)) *dst = t;
            __syncwarp();
            after(lane);
        }
        __syncthreads();
    }
    template <class LF, class FF> __device__ __forceinline__ void warp_fit(LF lane_f, FF final_f) {
        if (threadIdx.x < 32) {
            FitSums f = lane_f(int(threadIdx.x));
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                f.s0 += __shfl_xor_sync(0xffffffffu, f.s0, o);
                f.su += __shfl_xor_sync(0xffffffffu, f.su, o);
                f.sv += __shfl_xor_sync(0xffffffffu, f.sv, o);
                f.suv += __shfl_xor_sync(0xffffffffu, f.suv, o);
                f.suu += __shfl_xor_sync(0xffffffffu, f.suu, o);
                f.svv += __shfl_xor_sync(0xffffffffu, f.svv, o);
            }
            if (threadIdx.x == 0) final_f(f);
        }
        __syncthreads();
    }
    // ---- staged tiles (C::STAGE): TMA -> shared memory, completion on an mbarrier --------
    unsigned mbar, tiles;            // shared-memory addresses
    unsigned parity;                 // phase of the current pair's tiles
    const CcParams *prm;
    int next_pair;                   // pair whose tiles tiles_release() stages (-1: none)

    __device__ __forceinline__ void tiles_issue(int p) {        // one thread
        if (prm->boxes) {
            // frame mode: five boxes of (NMAX + 4) x NMAX pixels starting at column x1 & ~3
            constexpr unsigned bytes = unsigned(C::TILE_BYTES);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                         :: "r"(mbar), "r"(bytes) : "memory");
            const int4 b = __ldg(reinterpret_cast<const int4 *>(prm->boxes) + p);
            const int xa = b.x & ~3;
#pragma unroll
            for (int s = 0; s < 5; ++s)
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
                             " [%0], [%1, {%2, %3}], [%4];"
                             :: "r"(tiles + s * unsigned(C::TILE_BYTES / 5)),
                                "l"(reinterpret_cast<unsigned long long>(prm->tmap + s)),
                                "r"(xa), "r"(b.y), "r"(mbar) : "memory");
        } else {
            constexpr unsigned bytes = unsigned(20 * C::NMAX * C::NMAX);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                         :: "r"(mbar), "r"(bytes) : "memory");
            const float *src = prm->src[0] + size_t(p) * prm->pair_stride[0];
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes"
                         " [%0], [%1], %2, [%3];"
                         :: "r"(tiles), "l"(src), "r"(bytes), "r"(mbar) : "memory");
        }
    }
    __device__ __forceinline__ void tiles_wait() {
        asm volatile("{\n\t.reg .pred P1;\n\t"
                     "SPX_WAIT_%=:\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
                     "@P1 bra SPX_DONE_%=;\n\t"
                     "bra SPX_WAIT_%=;\n\t"
                     "SPX_DONE_%=:\n\t}"
                     :: "r"(mbar), "r"(parity) : "memory");
    }
    __device__ __forceinline__ void tiles_release() {
        if (threadIdx.x == 0 && next_pair >= 0) tiles_issue(next_pair);
    }
    __device__ __forceinline__ void cluster_sync() { cg::this_cluster().sync(); }
    __device__ __forceinline__ void cluster_arrive() { cg::this_cluster().barrier_arrive(); }
    __device__ __forceinline__ void cluster_wait() { cg::this_cluster().barrier_wait(); }
    __device__ __forceinline__ void global_fence() { __threadfence(); }
    __device__ __forceinline__ void or_status(int *p, int v) { atomicOr(p, v); }
    __device__ __forceinline__ float load_icc(const float *p) { return __ldcg(p); }
};

// Pull the five tiles of a (future) pair into L2 so that its loads do not pay
// the HBM latency: one prefetch per 128-byte line, spread over the CTA.
// The lane -> (row within a group, tile, line) mapping depends only on nx, so
// it is computed once per kernel (PfMap) and the per-pair loop has no divisions.
struct PfMap { int s, seg, rr, rpi; };      // s < 0: lane idle; rpi: rows per warp iteration
__device__ __forceinline__ PfMap prefetch_map(int nx, int lane) {
    PfMap m;
    const int segs = (nx * 4 + 127) / 128 + 1;           // lines a row can touch (unaligned start)
    const int per = 5 * segs;
    m.rpi = per <= 32 ? 32 / per : 0;                    // 0: generic path (rows wider than 160 px)
    m.rr = m.rpi ? lane / per : 0;
    const int k = lane - m.rr * per;
    m.s = k / segs; m.seg = k - m.s * segs;
    if (m.rr >= m.rpi) m.s = -1;
    return m;
}
template <class C>
__device__ __forceinline__ void prefetch_pair(const CcParams &prm, int k, const PfMap &m) {
    if (k >= prm.n) return;
    if (prm.shapes) {                        // ragged stack: the pair is one contiguous block
        const int p = prm.order ? __ldg(prm.order + k) : k;
        const long long elems = 5LL * __ldg(prm.shapes + 2 * p) * __ldg(prm.shapes + 2 * p + 1);
        const uintptr_t lo = reinterpret_cast<uintptr_t>(prm.src[0] + __ldg(prm.offsets + p));
        const uintptr_t hi = lo + 4 * elems;
        for (uintptr_t line = (lo & ~uintptr_t(127)) + 128 * threadIdx.x; line < hi; line += 128 * C::T)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(line));
        return;
    }
    const int p = k;
    int x1 = 0, y1 = 0;
    if (prm.boxes) {
        const int4 b = __ldg(reinterpret_cast<const int4 *>(prm.boxes) + p);
        x1 = b.x; y1 = b.y;
    }
    // columns / rows of the box that exist in the source image
    const int xa = x1 < 0 ? 0 : x1, xb = x1 + prm.nx > prm.W ? prm.W : x1 + prm.nx;
    const int ya = y1 < 0 ? 0 : y1, yb = y1 + prm.ny > prm.H ? prm.H : y1 + prm.ny;
    if (xa >= xb || ya >= yb) return;
    const int rows = yb - ya;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (m.rpi) {
        if (m.s < 0) return;
        const float *base = prm.src[m.s] + size_t(p) * prm.pair_stride[m.s] + (long long)ya * prm.pitch;
        for (int r = warp * m.rpi + m.rr; r < rows; r += C::NW * m.rpi) {
            const float *row = base + (long long)r * prm.pitch;
            const uintptr_t line = (reinterpret_cast<uintptr_t>(row + xa) & ~uintptr_t(127)) + m.seg * 128;
            if (line < reinterpret_cast<uintptr_t>(row + xb))
                asm volatile("prefetch.global.L2 [%0];" ::"l"(line));
        }
    } else {
        const int segs = (prm.nx * 4 + 127) / 128 + 1;
        for (int s = 0; s < 5; ++s) {
            const float *base = prm.src[s] + size_t(p) * prm.pair_stride[s] + (long long)ya * prm.pitch;
            for (int r = warp; r < rows; r += C::NW) {
                const float *row = base + (long long)r * prm.pitch;
                for (int seg = lane; seg < segs; seg += 32) {
                    const uintptr_t line = (reinterpret_cast<uintptr_t>(row + xa) & ~uintptr_t(127)) + seg * 128;
                    if (line < reinterpret_cast<uintptr_t>(row + xb))
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(line));
                }
            }
        }
    }
}

template <class C>
__global__ void __launch_bounds__(C::T, C::MINB) cc_fit_kernel(const __grid_constant__ CcParams prm) {
    extern __shared__ __align__(128) unsigned char smem[];
    using L = Layout<C>;
    uint16_t *rrow = reinterpret_cast<uint16_t *>(smem + L::off_rrow);
    // one pair per CTA, or per cluster of CL CTAs
    const int rank = C::CL > 1 ? int(cg::this_cluster().block_rank()) : 0;
    const int unit = blockIdx.x / C::CL, nunits = gridDim.x / C::CL;
    unsigned char *slab = prm.workspace + size_t(unit) * prm.slab_bytes;

    Ctx<C> c;
    c.S = reinterpret_cast<cf *>(smem + L::off_S);
    c.W = C::W_SMEM ? reinterpret_cast<cf *>(smem + L::off_W)
                    : reinterpret_cast<cf *>(slab + prm.off_w);
    c.R = C::R_SMEM ? reinterpret_cast<cf *>(smem + L::off_R)
                    : reinterpret_cast<cf *>(slab + prm.off_r);
    c.rrow = rrow;
    c.ps = reinterpret_cast<PairState *>(smem + L::off_ps);
    c.rank = rank;
    c.col_lo = rank * C::NXL;
    c.xl = reinterpret_cast<Xchg *>(smem + L::off_xchg);
    c.win = reinterpret_cast<FitWin *>(smem + L::off_win);

    DevExec<C> ex;
    ex.red = reinterpret_cast<double *>(smem + L::off_red);
    if (C::STAGE) {
        ex.mbar = unsigned(__cvta_generic_to_shared(smem + L::off_mbar));
        ex.tiles = unsigned(__cvta_generic_to_shared(smem + L::off_tiles));
        ex.prm = &prm;
        ex.parity = 0u;
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(ex.mbar) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            if (unit < prm.n) ex.tiles_issue(unit);              // the first pair of this CTA
        }
    }

    __shared__ int s_nrrow;
    if (threadIdx.x == 0) {
        s_nrrow = build_rrow<C>(rrow, rank);
        for (int k = 0; k < 4; ++k) (&c.xl->det[0][0])[k] = 0u;
    }
    __syncthreads();
    c.n_rrow = s_nrrow;
    if (C::CL > 1) cg::this_cluster().sync();      // the peers are resident: DSMEM may be used

    const size_t icc_elems = size_t(4) * prm.ny * prm.nx;
    const PfMap pfmap = prefetch_map(prm.nx, threadIdx.x & 31);
    c.dslot = 0;
    // Pair k of iteration i: the first two are static (unit, unit + nunits); with a claim
    // counter thread 0 claims the pair of iteration i + 2 at the start of iteration i and
    // parks it in s_claim[i & 1], which everybody reads after the barriers of the pair.
    __shared__ int s_claim[2];
    // (cluster classes: rank 0 claims for the cluster and writes the pair into every CTA's
    // exchange block; the cluster barriers of the pair order the write before the reads)
    const bool dynamic = !C::STAGE && prm.claim != nullptr;
    int knext = unit + nunits;
    for (int k = unit, it = 0; k < prm.n; ++it, c.dslot ^= 1) {
        if (dynamic && threadIdx.x == 0 && rank == 0) {
            const int v = 2 * nunits + int(atomicAdd(prm.claim, 1u));
            if (C::CL == 1) s_claim[it & 1] = v;
            else for (int r = 0; r < C::CL; ++r) X_of<C>(c, r)->claim[it & 1] = v;
        }
        const int p = prm.order ? __ldg(prm.order + k) : k;
        PairJob job;
        job.x1 = 0; job.y1 = 0; job.status0 = 0;
        job.ny = prm.ny; job.nx = prm.nx;
        job.pitch = prm.pitch; job.H = prm.H; job.W = prm.W;
        size_t icc_off = size_t(p) * icc_elems;
        if (prm.shapes) {
            job.ny = __ldg(prm.shapes + 2 * p); job.nx = __ldg(prm.shapes + 2 * p + 1);
            // a shape outside this launch's size class cannot be processed: flag it and
            // run a minimal pair instead (the host buckets by class, so this is a caller bug)
            if (job.ny < SPX_MIN_CUTOUT || job.nx < SPX_MIN_CUTOUT || job.ny > C::NMAX || job.nx > C::NMAX) {
                job.status0 = SPX_S_BOX_MISMATCH;
                job.ny = job.nx = SPX_MIN_CUTOUT;
            }
            const size_t tile = size_t(job.ny) * job.nx;
            const float *b0 = prm.src[0] + __ldg(prm.offsets + p);
            for (int s = 0; s < 5; ++s) job.base[s] = b0 + s * tile;
            job.pitch = job.nx; job.H = job.ny; job.W = job.nx;
            if (prm.icc_offsets) icc_off = size_t(__ldg(prm.icc_offsets + p));
        } else {
            if (prm.boxes) {
                const int4 b = __ldg(reinterpret_cast<const int4 *>(prm.boxes) + p);
                job.x1 = b.x; job.y1 = b.y;
                if (b.z - b.x + 1 != prm.nx || b.w - b.y + 1 != prm.ny)
                    job.status0 = SPX_S_BOX_MISMATCH;
            }
            for (int s = 0; s < 5; ++s) job.base[s] = prm.src[s] + size_t(p) * prm.pair_stride[s];
        }
        if (C::STAGE) {
            // the tiles are (being) staged in shared memory: a full-size cutout at the origin
            // of a NMAX x NMAX "image" (pixels outside the frame arrived as zeros)
            const float *t0 = reinterpret_cast<const float *>(smem + L::off_tiles);
            const int pitch = prm.boxes ? C::STAGE_PITCH : C::NMAX;
            for (int s = 0; s < 5; ++s) job.base[s] = t0 + s * (C::NMAX * pitch);
            job.x1 = prm.boxes ? (job.x1 & 3) : 0;          // the staged box starts at x1 & ~3
            job.y1 = 0; job.pitch = pitch; job.H = C::NMAX; job.W = pitch;
            ex.next_pair = knext < prm.n ? knext : -1;
        }
        c.icc = prm.icc_out ? prm.icc_out + icc_off
                            : reinterpret_cast<float *>(slab + prm.off_icc);
        job.cc_type = prm.cc_type;
        job.flags = prm.flags; job.patch_icc = prm.icc_out != nullptr;
        if (rank == 0 && !C::STAGE) prefetch_pair<C>(prm, knext, pfmap);
        // peak == NULL: packed [n,4] float64 result table (dx, dy, peak, status).  ONE call
        // site: a second inlined copy of process_pair costs 11 % (instruction cache)
        const bool packed = prm.peak == nullptr;
        process_pair<C>(c, ex, job, prm.shifts + (packed ? 4 : 2) * size_t(p),
                        packed ? nullptr : prm.peak + p, packed ? nullptr : prm.status + p);
        if (C::STAGE) ex.parity ^= 1u;
        k = knext;
        knext = !dynamic ? knext + nunits : (C::CL == 1 ? s_claim[it & 1] : c.xl->claim[it & 1]);
    }
    if (C::CL > 1) cg::this_cluster().sync();      // do not exit while a peer may still read my W / R
}

// ---- host side ------------------------------------------------------------------
struct SlabPlan { size_t off_icc, off_w, off_r, bytes; int max_ctas_per_sm, cluster; };

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

template <class C>
static SlabPlan plan_slab(int ny, int nx, bool want_icc) {
    SlabPlan s;
    size_t o = 0;
    // surface scratch: only when nobody asked for the surface AND it cannot stay in W
    s.off_icc = o; if (!want_icc && !SurfaceInW<C>::value) o += align_up(size_t(16) * ny * nx, 256);
    s.off_w = o;   if (!C::W_SMEM) o += align_up(C::W_ELEMS * 8, 256);
    s.off_r = o;   if (!C::R_SMEM) o += align_up(C::R_ELEMS * 8, 256);
    s.bytes = align_up(o, 256);
    s.max_ctas_per_sm = C::MINB;
    s.cluster = C::CL;
    return s;
}

static int num_sms() {
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) {
        cudaGetLastError();
        sms = 148;                       // B200
    }
    return sms;
}

static int pick_cfg(int ny, int nx) {
    const int m = ny > nx ? ny : nx;
    if (m <= 20) return 20;
    if (m <= 32) return 32;
    if (m <= 42) return 42;
    if (m <= 64) return 64;
    if (m <= 84) return 84;
    if (m <= 128) return 128;
    if (m <= 170) return 170;
    if (m <= 256) return 256;
    return 0;
}

// ---- TMA staging ------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                  const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
        else
            cudaGetLastError();
    }
    return fn;
}

// Can this launch use the TMA-staged form of its size class?  Uniform full-size cutouts only;
// stack mode needs the five tiles of a pair contiguous and 16-byte aligned (one bulk copy),
// frame mode a 16-byte aligned base and row pitch (tensor maps).  SPX_STAGE=1 opts in.
// Descriptor slots in global memory: a small ring per device (allocated once), so that
// launches on different streams do not share a slot while one of them is still running.
static CUtensorMap *tmap_slot(cudaStream_t st, const CUtensorMap (&host)[5]) {
    constexpr int kSlots = 64;
    static CUtensorMap *ring[64] = {nullptr};
    static unsigned next[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    if (!ring[dev] && cudaMalloc(&ring[dev], kSlots * 5 * sizeof(CUtensorMap)) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    CUtensorMap *slot = ring[dev] + 5 * (next[dev]++ % kSlots);
    if (cudaMemcpyAsync(slot, host, 5 * sizeof(CUtensorMap), cudaMemcpyHostToDevice, st) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return slot;
}

static bool use_stage(CcParams &prm, int cfg, cudaStream_t st) {
    static int want = -1;
    if (want < 0) { const char *e = getenv("SPX_STAGE"); want = (e && atoi(e) != 0) ? 1 : 0; }
    if (!want || (cfg != 32 && cfg != 64) || prm.shapes || prm.ny != cfg || prm.nx != cfg) return false;
    const size_t tile = size_t(cfg) * cfg;
    if (!prm.boxes) {
        for (int s = 0; s < 5; ++s)
            if (prm.src[s] != prm.src[0] + s * tile || prm.pair_stride[s] != prm.pair_stride[0]) return false;
        return reinterpret_cast<uintptr_t>(prm.src[0]) % 16 == 0 && (prm.pair_stride[0] * 4) % 16 == 0;
    }
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc || (prm.pitch * 4) % 16 != 0) return false;
    alignas(64) CUtensorMap host[5];
    for (int s = 0; s < 5; ++s) {
        if (reinterpret_cast<uintptr_t>(prm.src[s]) % 16 != 0) return false;
        const cuuint64_t dims[2] = {cuuint64_t(prm.W), cuuint64_t(prm.H)};
        const cuuint64_t strides[1] = {cuuint64_t(prm.pitch) * 4};
        const cuuint32_t box[2] = {cuuint32_t(cfg + 4), cuuint32_t(cfg)}, estr[2] = {1, 1};
        if (enc(&host[s], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(prm.src[s]), dims,
                strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return false;
    }
    prm.tmap = tmap_slot(st, host);
    return prm.tmap != nullptr;
}

// Claim counters of the dynamic work distribution: a ring of device words (no allocation);
// every launch takes the next one and zeroes it in stream order.
__device__ unsigned g_claim_ring[256];
static unsigned *claim_slot(cudaStream_t st) {
    static unsigned *base[64] = {nullptr};
    static unsigned next[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    if (!base[dev]) {
        void *p = nullptr;
        if (cudaGetSymbolAddress(&p, g_claim_ring) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        base[dev] = static_cast<unsigned *>(p);
    }
    unsigned *slot = base[dev] + (next[dev]++ % 256);
    if (cudaMemsetAsync(slot, 0, sizeof(unsigned), st) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return slot;
}

static SlabPlan plan_for(int cfg, int ny, int nx, bool want_icc) {
    switch (cfg) {
        case 20: return plan_slab<Cfg20>(ny, nx, want_icc);
        case 42: return plan_slab<Cfg42>(ny, nx, want_icc);
        case 84: return plan_slab<Cfg84>(ny, nx, want_icc);
        case 170: return plan_slab<Cfg170>(ny, nx, want_icc);
        case 32: return plan_slab<Cfg32>(ny, nx, want_icc);
        case 64: return plan_slab<Cfg64>(ny, nx, want_icc);
        case 128: return plan_slab<Cfg128>(ny, nx, want_icc);
        default: return plan_slab<Cfg256>(ny, nx, want_icc);
    }
}

// number of pairs in flight = CTAs (or clusters) in the persistent grid
static int units_for(int n, const SlabPlan &sp) {
    const long long cap = (long long)num_sms() * sp.max_ctas_per_sm / sp.cluster;
    return int(n < cap ? n : cap);
}

template <class C>
static int launch(CcParams &prm, const SlabPlan &sp, cudaStream_t stream) {
    const size_t smem = Layout<C>::total;
    cudaError_t e = cudaFuncSetAttribute(cc_fit_kernel<C>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
    if (e != cudaSuccess) return int(e);
    // persistent grid of THIS configuration (a staged form runs fewer CTAs per SM than the
    // plain form the workspace was sized for)
    const long long cap = (long long)num_sms() * C::MINB / C::CL;
    int grid = int(prm.n < cap ? prm.n : cap) * C::CL;
    (void)sp;
    if (C::CL == 1) {
        cc_fit_kernel<C><<<grid, C::T, smem, stream>>>(prm);
    } else {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(C::T);
        cfg.dynamicSmemBytes = smem; cfg.stream = stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = C::CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        // persistent grid: never more clusters than can be co-resident (GPCs with an odd
        // number of usable SMs cannot pair all of them), or the tail runs as a second wave
        static int max_clusters[64] = {0};
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev >= 0 && dev < 64) {
            if (max_clusters[dev] == 0) {
                int mc = 0;
                cudaLaunchConfig_t q = cfg;
                q.gridDim = dim3(num_sms() * C::MINB / C::CL * C::CL);      // whole clusters
                if (cudaOccupancyMaxActiveClusters(&mc, cc_fit_kernel<C>, &q) != cudaSuccess || mc <= 0) {
                    cudaGetLastError();
                    mc = -1;
                }
                max_clusters[dev] = mc;
                if (getenv("SPX_DEBUG")) fprintf(stderr, "subpixal_b200: max active clusters = %d\n", mc);
            }
            if (max_clusters[dev] > 0 && grid > max_clusters[dev] * C::CL) grid = max_clusters[dev] * C::CL;
        }
        cfg.gridDim = dim3(grid);
        e = cudaLaunchKernelEx(&cfg, cc_fit_kernel<C>, prm);
        if (e != cudaSuccess) return int(e);
    }
    e = cudaGetLastError();
    return e == cudaSuccess ? 0 : int(e);
}

static int run_cc(CcParams &prm, void *workspace, size_t workspace_bytes, void *stream) {
    // peak and status both NULL: `shifts` is the packed [n,4] float64 table
    if (prm.n < 0 || !prm.shifts || (!prm.peak) != (!prm.status)) return SPX_E_BADARG;
    for (int s = 0; s < 5; ++s) if (!prm.src[s]) return SPX_E_BADARG;
    if (prm.ny < SPX_MIN_CUTOUT || prm.nx < SPX_MIN_CUTOUT) return SPX_E_UNSUPPORTED;
    const int cfg = pick_cfg(prm.ny, prm.nx);
    if (!cfg) return SPX_E_UNSUPPORTED;
    if (prm.n == 0) return 0;
    if (prm.cc_type != SPX_NCC && prm.cc_type != SPX_ZNCC) prm.cc_type = SPX_CC;  // cc.py:108
    const SlabPlan sp = plan_for(cfg, prm.ny, prm.nx, prm.icc_out != nullptr);
    const size_t need = sp.bytes * size_t(units_for(prm.n, sp));
    if (need > 0 && (!workspace || workspace_bytes < need)) return SPX_E_WORKSPACE;
    prm.workspace = static_cast<unsigned char *>(workspace);
    prm.slab_bytes = sp.bytes;
    prm.off_icc = sp.off_icc; prm.off_w = sp.off_w; prm.off_r = sp.off_r;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    static int want_dynamic = -1;       // SPX_STATIC=1: static round-robin assignment (A/B knob)
    if (want_dynamic < 0) { const char *e = getenv("SPX_STATIC"); want_dynamic = (e && atoi(e) != 0) ? 0 : 1; }
    prm.claim = want_dynamic ? claim_slot(st) : nullptr;     // (cluster classes ignore it)
#ifdef SPX_ONLY          // developer knob: compile ONE size class (fast ptxas / SASS turn-around)
#define SPX_CASE(N, CFG) case N: if (SPX_ONLY == N) return launch<CFG>(prm, sp, st); return SPX_E_UNSUPPORTED;
    switch (cfg) {
        SPX_CASE(20, Cfg20) SPX_CASE(42, Cfg42) SPX_CASE(84, Cfg84) SPX_CASE(170, Cfg170)
        SPX_CASE(32, Cfg32) SPX_CASE(64, Cfg64) SPX_CASE(128, Cfg128) SPX_CASE(256, Cfg256)
        default: return SPX_E_UNSUPPORTED;
    }
#undef SPX_CASE
#else
    if (use_stage(prm, cfg, st))
        return cfg == 32 ? launch<Cfg32S>(prm, sp, st) : launch<Cfg64S>(prm, sp, st);
    switch (cfg) {
        case 20: return launch<Cfg20>(prm, sp, st);
        case 42: return launch<Cfg42>(prm, sp, st);
        case 84: return launch<Cfg84>(prm, sp, st);
        case 170: return launch<Cfg170>(prm, sp, st);
        case 32: return launch<Cfg32>(prm, sp, st);
        case 64: return launch<Cfg64>(prm, sp, st);
        case 128: return launch<Cfg128>(prm, sp, st);
        default: return launch<Cfg256>(prm, sp, st);
    }
#endif
}

}  // namespace spx

extern "C" {

size_t spx_cc_workspace_bytes(int n, int ny, int nx, int want_icc) {
    const int cfg = spx::pick_cfg(ny, nx);
    if (!cfg || n <= 0 || ny < 1 || nx < 1) return 0;
    const spx::SlabPlan sp = spx::plan_for(cfg, ny, nx, want_icc != 0);
    return sp.bytes * size_t(spx::units_for(n, sp));
}

int spx_cc_fit_stack(const float *ref, const float *im00, const float *im10,
                     const float *im01, const float *im11,
                     const int64_t pair_stride_host[5],
                     int n, int ny, int nx, int cc_type,
                     double *shifts, int32_t *peak, int32_t *status,
                     float *icc, void *workspace, size_t workspace_bytes,
                     void *stream) {
    if (!pair_stride_host) return SPX_E_BADARG;
    spx::CcParams prm{};
    const float *src[5] = {ref, im00, im10, im01, im11};
    for (int s = 0; s < 5; ++s) { prm.src[s] = src[s]; prm.pair_stride[s] = pair_stride_host[s]; }
    prm.pitch = nx; prm.H = ny; prm.W = nx;
    prm.boxes = nullptr;
    prm.n = n; prm.ny = ny; prm.nx = nx; prm.cc_type = cc_type; prm.flags = 0;
    prm.shifts = shifts; prm.peak = peak; prm.status = status; prm.icc_out = icc;
    return spx::run_cc(prm, workspace, workspace_bytes, stream);
}

int spx_cc_fit_ragged(const float *stack, const int64_t *offsets, const int32_t *shapes,
                      const int32_t *order, int n, int max_ny, int max_nx, int cc_type,
                      double *shifts, int32_t *peak, int32_t *status,
                      float *icc, const int64_t *icc_offsets,
                      void *workspace, size_t workspace_bytes, void *stream) {
    if (!stack || !offsets || !shapes || (icc && !icc_offsets)) return SPX_E_BADARG;
    spx::CcParams prm{};
    for (int s = 0; s < 5; ++s) { prm.src[s] = stack; prm.pair_stride[s] = 0; }
    prm.pitch = max_nx; prm.H = max_ny; prm.W = max_nx;
    prm.boxes = nullptr;
    prm.shapes = shapes; prm.offsets = reinterpret_cast<const long long *>(offsets);
    prm.order = order; prm.icc_offsets = reinterpret_cast<const long long *>(icc_offsets);
    prm.n = n; prm.ny = max_ny; prm.nx = max_nx; prm.cc_type = cc_type; prm.flags = 0;
    prm.shifts = shifts; prm.peak = peak; prm.status = status; prm.icc_out = icc;
    return spx::run_cc(prm, workspace, workspace_bytes, stream);
}

int spx_cc_fit_frames(const float *const frames_host[5],
                      int frame_h, int frame_w, int64_t frame_pitch,
                      const int32_t *boxes, int n, int ny, int nx,
                      int cc_type, unsigned flags,
                      double *shifts, int32_t *peak, int32_t *status,
                      float *icc, void *workspace, size_t workspace_bytes,
                      void *stream) {
    if (!frames_host || !boxes || frame_h < 1 || frame_w < 1 || frame_pitch < frame_w)
        return SPX_E_BADARG;
    if (int64_t(frame_h) * frame_pitch >= (int64_t(1) << 31)) return SPX_E_UNSUPPORTED;
    spx::CcParams prm{};
    for (int s = 0; s < 5; ++s) { prm.src[s] = frames_host[s]; prm.pair_stride[s] = 0; }
    prm.pitch = frame_pitch; prm.H = frame_h; prm.W = frame_w;
    prm.boxes = boxes;
    prm.n = n; prm.ny = ny; prm.nx = nx; prm.cc_type = cc_type; prm.flags = flags;
    prm.shifts = shifts; prm.peak = peak; prm.status = status; prm.icc_out = icc;
    return spx::run_cc(prm, workspace, workspace_bytes, stream);
}

}  // extern "C"
