// Cross-correlation engine: the per-thread work of every pass of the fused
// gather -> normalise -> 4-phase correlation -> interlace -> argmax pipeline.
//
// Mathematics (reference cc.py:98-128):
//   C_ab(ly,lx) = sum_{j,i} ref[j,i] * im_ab[j+ly, i+lx]
//   ICC[2k+b, 2l+a] = C_ab(k - yc, l - xc),  yc=(ny-1)//2, xc=(nx-1)//2
// computed as a circular correlation of size P x P, P >= n + max(c, n-1-c)
// (alias-free for the kept lags), with the image tiles placed at offset
// (yc, xc) so that output sample m IS lag m - c:
//   c = IFFT2( conj(FFT2(ref)) * FFT2(im00 + i*im10) )  ->  Re = C_00, Im = C_10
// (two real phases ride on one complex transform; ref is real so linearity
// keeps them separate).  The four phase surfaces are written interlaced.
//
// Data flow per phase pair (1-D FFT of length P = E1*E2 done as two register
// codelets with one exchange through memory; columns first so that lanes run
// along x for the dense input tiles):
//   C1  tiles -> E1-point FFT down columns (+twiddle)          -> W[P][nx]
//   C2  E2-point FFT down columns, in place                    -> W (ky scrambled)
//   for each batch of 32 rows (lanes = rows):
//     X1  E1-point FFT along x of the nx non-zero samples      -> S[32][P]
//     X2  E2-point FFT, * conj(R), E2-point inverse FFT        -> S
//     X3  E1-point inverse FFT, keep the nx wanted lags        -> W[P][nx]
//   IC1 E2-point inverse FFT down columns (+twiddle), in place -> W
//   IC2 E1-point inverse FFT, keep ny lags, scale, interlace   -> ICC, argmax
// The reference image goes through C1, C2, X1 and a final E2-point FFT whose
// conjugate, times the normalisation scale, is stored as R (half spectrum,
// ky in [0, P/2], Hermitian symmetry supplies the rest).
//
// Slab classes (170 / 256 px: W and R in a per-CTA global slab, C::FUSE): the column
// passes run per 32-column chunk through the scratch S -
//   C1  tiles -> chunk [P][32] in S;   C2  chunk -> W slab (256-byte row segments)
//   row batches of 32 rows as above (a batch = k1 group g with group E1 - g, so the two
//   rows that read one R row share it; 128-bit copies between the slab and S)
//   IC1 W slab -> chunk in S;          IC2 chunk -> ICC, argmax
// - so W makes one trip to the slab per direction change.
#pragma once
#include <limits.h>
#if !defined(SPX_HOST_EMU)
#include <cooperative_groups.h>
#endif
#include "spx_common.cuh"
#include "spx_fft.cuh"
#include "spx_peak.cuh"

// item loops of the passes: optional unrolling so that the loads of the next item
// overlap the arithmetic of the current one (knobs for tools/variant.sh)
#if defined(SPX_UNROLL_E2) && !defined(SPX_HOST_EMU)
#define SPX_UNROLL_SMALL _Pragma("unroll 2")
#else
#define SPX_UNROLL_SMALL
#endif
#if defined(SPX_UNROLL_E1) && !defined(SPX_HOST_EMU)
#define SPX_UNROLL_BIG _Pragma("unroll 2")
#else
#define SPX_UNROLL_BIG
#endif
#if !defined(SPX_HOST_EMU)
#define SPX_NO_UNROLL _Pragma("unroll 1")
#else
#define SPX_NO_UNROLL
#endif

namespace spx {

// --------------------------------------------------------------------------
// twiddle tables: __constant__ on the device, plain statics for the host emu
// --------------------------------------------------------------------------
// two-step twiddle W_P^(n2*k1) from the [n2][k1] table (one of the two indices is
// a compile-time constant at every call site -> base register + immediate)
#ifndef SPX_HOST_EMU
template <int P, int E1, bool INV> __constant__ TwTable2<P, E1, INV> c_tw2 = TwTable2<P, E1, INV>();
#endif
template <int P, int E1, bool INV> struct HostTw2 { static const TwTable2<P, E1, INV> t; };
template <int P, int E1, bool INV>
const TwTable2<P, E1, INV> HostTw2<P, E1, INV>::t = TwTable2<P, E1, INV>();
// forward twiddle W_P^(n2*k1) / its conjugate (inverse transforms)
template <class C, bool INV = false> SPX_HD cf tw2(int n2, int k1) {
#if defined(__CUDA_ARCH__) && !defined(SPX_HOST_EMU)
    return c_tw2<C::P, C::E1, INV>.w[n2 * C::E1 + k1];
#else
    return HostTw2<C::P, C::E1, INV>::t.w[n2 * C::E1 + k1];
#endif
}

// --------------------------------------------------------------------------
// configuration
// --------------------------------------------------------------------------
// INPLACE: W is P x (P+1) and the row stage (X1, X2, X3) works in place on whole
// rows - three barriers per row stage and no separate scratch.  Otherwise W is
// P x (NMAX+1) and the row stage goes through a 32-row scratch S, batch by batch
// (used when W lives in global memory).
//
// CL = 2, 4, 8: one pair is processed by a thread-block cluster of CL CTAs.  W is
// split by ROWS: the rows r = k1*E2 + k2 of one k1 group all have ky = k1 + E1*k2
// congruent to k1 modulo CL (CL divides E1).  CTA `rank` owns the groups with
// k1 % CL == rank: C2, the whole row stage and IC1 are then purely local.  The half
// spectrum R is split the same way - row q lives in CTA q % CL at local row q / CL:
// a W row with ky <= P/2 finds its R row (q = ky) at home, a row with ky > P/2 needs
// q = P - ky, which is congruent to -rank: the rows of CTA (CL - rank) % CL, read over
// DSMEM (for CL = 2 that is the CTA itself).  Besides those reads only C1 (it stores the
// outputs of the k1 groups it does not own into their owners' shared memory) and IC2
// (it loads them back) cross the cluster; the tile columns, and therefore all HBM
// traffic, are split evenly.
//
// STAGE: the five tiles of a pair are staged in shared memory by the TMA (one bulk copy of
// the contiguous [5][NMAX][NMAX] block of a stack, or five 2-D tensor copies out of the
// frames whose out-of-bounds zero fill IS Cutout's 'fill' mode, cutout.py:737-752 - measured
// on B200: rows / columns outside the tensor read as zeros, but a box whose first column is
// not on a 16-byte boundary faults with "illegal instruction"), signalled
// on an mbarrier; the copy for pair p+1 is issued as soon as the last C1 pass of pair p has
// read its tiles, so the statistics and C1 passes read shared memory and never wait for HBM.
// Only for full-size cutouts (ny = nx = NMAX); costs 20 NMAX^2 bytes of shared memory.
template <int P_, int E1_, int E2_, int NMAX_, int T_, int MINB_, bool WSM_, bool RSM_,
          bool INPLACE_ = false, int CL_ = 1, bool STAGE_ = false>
struct CcCfg {
    static constexpr int P = P_, E1 = E1_, E2 = E2_, NMAX = NMAX_, T = T_, MINB = MINB_;
    static constexpr bool W_SMEM = WSM_, R_SMEM = RSM_, INPLACE = INPLACE_;
    static constexpr int CL = CL_;
    static constexpr bool STAGE = STAGE_;
    // (frame mode stages NMAX + 4 columns per row: the TMA wants the first column of a box
    // on a 16-byte boundary, so the box starts at x1 & ~3 and the tile is read at x1 & 3)
    static constexpr int STAGE_PITCH = NMAX_ + 4;
    static constexpr size_t TILE_BYTES = STAGE_ ? size_t(20) * NMAX_ * STAGE_PITCH : 0;
    // the compile-time fast path for full-size interior cutouts places the tiles at
    // offset NMAX/2, which must be a multiple of E2 (see c1_pass)
    static constexpr bool HAS_FAST = (NMAX_ / 2) % E2_ == 0 && NMAX_ % CL_ == 0;
    static constexpr int NW = T / 32;
    static constexpr int NXL = (NMAX + CL - 1) / CL;   // tile columns loaded / finished by one CTA
    static constexpr int E1L = E1 / CL;      // k1 groups (of E2 rows each) held by one CTA
    static constexpr int PL = P / CL;        // rows of W held by one CTA
    // pitch of W (complex): odd in shared memory (bank conflicts); in a global slab a
    // multiple of 16 elements, so that rows and 32-column chunks start on 128-byte lines
#ifndef SPX_SLAB_ODD_PITCH
    static constexpr int WP = INPLACE ? P + 1 : (WSM_ ? NMAX + 1 : (NMAX + 15) / 16 * 16);
#else
    static constexpr int WP = INPLACE ? P + 1 : NMAX + 1;
#endif
    // how the row stage reaches its rows: in place in W / directly in W (shared
    // memory, lanes along rows) / staged through S by coalesced copies (W global)
    static constexpr bool DIRECT = !INPLACE && WSM_;
    // W in global memory (slab classes): the two column passes on either side of the row
    // stage are FUSED per 32-column chunk through the scratch S (C1 -> C2 -> copy out;
    // copy in -> IC1 -> IC2), so W makes one trip to the slab per direction change instead
    // of three (see phase_pair)
#ifndef SPX_NO_FUSE
    static constexpr bool FUSE = !WSM_ && !INPLACE_ && CL_ == 1;
#else
    static constexpr bool FUSE = false;
#endif
    static constexpr int CP = 32;            // pitch of a column chunk held in S: [PL][32]
    // cluster classes whose column split falls on 32-column chunks: SPLIT-PHASE cluster
    // barriers around the two column passes that need the peers' rows.  After C1 a CTA
    // arrives, runs C2 on the columns it loaded itself (written by nobody else) and only then
    // waits for the peers' stores; before IC2 it runs IC1 on the PEERS' columns first,
    // arrives, finishes its own columns and waits - half a pass of slack per barrier.
#ifndef SPX_NO_SPLIT_SYNC
    static constexpr bool SPLIT = CL_ > 1 && ((NMAX_ + CL_ - 1) / CL_) % 32 == 0;
#else
    static constexpr bool SPLIT = false;
#endif
    static constexpr int SP = P + 1;         // pitch of the row scratch, odd
#ifndef SPX_SLAB_ODD_PITCH
    static constexpr int RP = RSM_ ? P + 1 : P;   // pitch of R: odd in shared memory, whole lines in a slab
#else
    static constexpr int RP = P + 1;
#endif
    static constexpr int RH = P / 2 + 1;     // rows of R (half spectrum)
    static constexpr int RHL = (RH + CL - 1) / CL;   // ... held by one CTA
    static constexpr int BATCH = 32;         // rows per row-op batch (= lanes)
    static constexpr int NCHUNK = (PL + 31) / 32, RCHUNK = (RHL + 31) / 32;
    static constexpr size_t W_ELEMS = size_t(PL) * WP;
    static constexpr size_t S_ELEMS = INPLACE ? 0 : size_t(BATCH) * SP;
    static constexpr size_t R_ELEMS = size_t(RHL) * RP;
    static_assert(!STAGE_ || (CL_ == 1 && WSM_ && (NMAX_ / 2) % E2_ == 0), "staged tiles: one CTA, fast path");
    static_assert(CL == 1 || ((CL == 2 || CL == 4 || CL == 8) && WSM_ && RSM_ && INPLACE_ &&
                              E1 % CL == 0 && P % CL == 0),
                  "cluster mode: W and R in shared memory, CL divides E1 and P");

    static_assert(E1 * E2 == P, "P = E1*E2");
    static_assert(P % 2 == 0 && NMAX % 2 == 0, "even sizes");
    static_assert(T % 32 == 0, "whole warps");
    // alias-free for every n <= NMAX: P >= n + n/2
    static_assert(P >= NMAX + NMAX / 2, "transform too small");
};

// Per-pair state, written by thread 0 / the block reductions.
struct PairState {
    const float *base[5];
    long long pitch[5];
    int y0[5], x0[5];      // cutout origin inside the source image
    int H, W;              // source image size (ny, nx in stack mode)
    int ny, nx, yc, xc;
    int cc_type;
    unsigned flags;
    float mu, mu_ref;      // ZNCC means (0 otherwise)
    float rscale;          // normalisation folded into R
    int ylo, yhi, xlo, xhi; // cutout rows/cols that exist in the source image
    int interior;          // box entirely inside the source image
    int report;            // status bits of the tiles whose NaN/Inf are reported
    int status;
    int keep, in_w;        // surface kept in W (see SurfaceInW); the phase pair W holds now
    int ysh, cout;         // W row of lag row 0, W column of lag column 0
    Best best;             // running argmax over the interlaced surface
};

// Values the CTAs of a cluster hand to each other (each CTA writes slot [rank] of
// EVERY copy, then a cluster barrier, then reads its local copy).
constexpr int kMaxCluster = 8;
struct Xchg {
    // (det comes first: single-CTA configurations only allocate the first 64 bytes)
    // degeneracy detector of the statistics-free NCC path (lazy_decide): det[s][0] evidence
    // bits (1: phases, 2: reference), det[s][1] bits of the largest |pixel| seen; in cluster
    // mode both CTAs update both copies with atomics.  Two slots, used by alternate pairs:
    // a CTA clears the NEXT pair's slot between the first and the last cluster barrier of
    // the current pair - the only window in which the peer can neither still read it
    // (previous pair) nor already write it (next pair).
    unsigned det[2][2];
    int claim[2];          // cluster mode: the pair rank 0 claimed for iteration i + 2 (slot i & 1)
    double stat[kMaxCluster][8];
    float bestv[kMaxCluster];
    int besti[kMaxCluster];
    int status[kMaxCluster];
};

// Surface kept on chip.  When the caller does not ask for the surface (and one CTA holds
// W), IC2 writes its outputs back IN PLACE into the W slots it read - W row n, column l
// then holds the complex sample (C_a0, C_a1)(lag row n - ysh, lag column l) of the phase
// pair just finished - and only the argmax runs over them.  The 5x5 fit needs samples of
// BOTH phase pairs (even / odd surface rows): the second pair's are still in W at the end;
// of the first pair a 7x7 window of complex samples around ITS best sample is saved before
// the second pair overwrites W.  The four lattices of one source peak within a sample of
// each other, so the window almost always covers the fit box; when it does not (global
// best far from the first pair's best), the needed samples of the second pair are saved
// the same way and the first pair is transformed once more.
//
// Measured (B200, stack mode, A/B on one box): keeping the surface in W is 1.7 % slower at
// 64 px and 4 % slower at 20 / 32 px than writing it to the per-CTA slab in global memory,
// which stays L2 resident (stores carry an evict-last policy): shared-memory bandwidth is
// what binds these kernels, the L2 has slack.  So the slab is the default and
// -DSPX_KEEP_SURFACE=1 builds the on-chip form (tests/test_emulator.py keeps it alive).
#if defined(SPX_KEEP_SURFACE) && SPX_KEEP_SURFACE
template <class C> struct SurfaceInW { static constexpr bool value = C::W_SMEM && C::CL == 1 && !C::STAGE; };
#else
template <class C> struct SurfaceInW { static constexpr bool value = false; };
#endif
constexpr int kWinR = 3, kWinN = 2 * kWinR + 1;
struct FitWin {
    cf v[2][kWinN][kWinN];     // [phase pair][lattice row - k0][lattice column - l0]
    int k0[2], l0[2];
};

template <class C>
struct Ctx {
    cf *W, *S, *R;
    float *icc;            // [2ny][2nx] interlaced surface of this pair
    const uint16_t *rrow;  // my rows of W (local index) with ky <= P/2
    PairState *ps;
    // cluster mode (CL > 1); rank = 0 and col_lo = 0 otherwise
    int rank, col_lo, n_rrow;
    int dslot;             // detector slot of the current pair (alternates, see Xchg)
    Xchg *xl;              // local exchange block
    FitWin *win;           // fit-sample windows (SurfaceInW configurations)
#ifdef SPX_HOST_EMU        // the emulator keeps a table of every CTA's buffers
    cf *Wall[kMaxCluster], *Rall[kMaxCluster];
    Xchg *xall[kMaxCluster];
#endif
};

// The same shared-memory object in CTA `r` of the cluster (distributed shared memory).
template <class C, class T> SPX_HD T *dsmem_peer(T *local, int r) {
#if defined(__CUDA_ARCH__) && !defined(SPX_HOST_EMU)
    if (C::CL == 1) return local;
    return cooperative_groups::this_cluster().map_shared_rank(local, r);
#else
    (void)r;
    return local;
#endif
}
template <class C> SPX_HD cf *W_of(const Ctx<C> &c, int r) {
#ifdef SPX_HOST_EMU
    return C::CL == 1 ? c.W : c.Wall[r];
#else
    return dsmem_peer<C>(c.W, r);
#endif
}
template <class C> SPX_HD cf *R_of(const Ctx<C> &c, int r) {
#ifdef SPX_HOST_EMU
    return C::CL == 1 ? c.R : c.Rall[r];
#else
    return dsmem_peer<C>(c.R, r);
#endif
}
template <class C> SPX_HD Xchg *X_of(const Ctx<C> &c, int r) {
#ifdef SPX_HOST_EMU
    return C::CL == 1 ? c.xl : c.xall[r];
#else
    return dsmem_peer<C>(c.xl, r);
#endif
}
// W base pointers of all CTAs of the cluster, indexed by compile-time k1 % CL
template <class C> struct WBases { cf *p[C::CL]; };
template <class C> SPX_HD WBases<C> w_bases(const Ctx<C> &c) {
    WBases<C> b;
    static_for<0, C::CL>([&](auto Rk) {
        constexpr int r = decltype(Rk)::value;
        b.p[r] = W_of<C>(c, r);
    });
    return b;
}

// number of tile columns held by this CTA
template <class C>
SPX_HD int local_cols(const Ctx<C> &c, int nx) {
    if (C::CL == 1) return nx;
    int n = nx - c.col_lo;
    return n < 0 ? 0 : (n > C::NXL ? C::NXL : n);
}

// Warp index as a value the compiler KNOWS is warp-uniform (so that n2 / k1,
// the predicates and the twiddle addresses derived from it live in uniform
// registers): broadcast lane 0's copy.
SPX_HD int warp_of(int tid) {
#if defined(__CUDA_ARCH__) && !defined(SPX_HOST_EMU)
    return __shfl_sync(0xffffffffu, tid >> 5, 0);
#else
    return tid >> 5;
#endif
}

// tile loads of a configuration: staged tiles live in shared memory (plain loads), tiles in
// global memory go through the read-only / streaming paths below
template <class C> SPX_HD float ld_tile(const float *p);
template <class C> SPX_HD float ld_tile_last(const float *p);
SPX_HD float ld_f32(const float *p) {
#if defined(__CUDA_ARCH__) && !defined(SPX_HOST_EMU)
    return __ldg(p);
#else
    return *p;
#endif
}
// last use of a tile pixel: streaming load (evict-first), so the input stream
// does not push the L2-resident per-CTA scratch out to HBM
SPX_HD float ld_f32_last(const float *p) {
#if defined(__CUDA_ARCH__) && !defined(SPX_HOST_EMU)
    return __ldcs(p);
#else
    return *p;
#endif
}

template <class C> SPX_HD float ld_tile(const float *p) { return C::STAGE ? *p : ld_f32(p); }
template <class C> SPX_HD float ld_tile_last(const float *p) { return C::STAGE ? *p : ld_f32_last(p); }

// Row r = k1*E2 + k2 holds frequency ky = k1 + E1*k2 (in-place two-step order).
// Cluster mode: CTA `rank` stores the groups k1 = 2j + rank at local rows j*E2 + k2.
template <class C> SPX_HD int ky_of_row(int lr, int rank) {
    return (lr / C::E2) * C::CL + rank + C::E1 * (lr % C::E2);
}


SPX_HD unsigned f32_bits(float v) { union { float f; unsigned u; } c; c.f = v; return c.u; }
SPX_HD float bits_f32(unsigned u) { union { float f; unsigned u; } c; c.u = u; return c.f; }
// order-preserving map float -> unsigned (so that integer atomicMin / atomicMax order floats)
SPX_HD unsigned fkey(float v) {
    const unsigned u = f32_bits(v);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
SPX_HD float fkey_inv(unsigned k) { return bits_f32((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k); }
SPX_HD void key_min(unsigned *p, unsigned v) {
#if defined(__CUDA_ARCH__) && !defined(SPX_HOST_EMU)
    atomicMin(p, v);
#else
    if (v < *p) *p = v;
#endif
}
SPX_HD void key_max(unsigned *p, unsigned v) {
#if defined(__CUDA_ARCH__) && !defined(SPX_HOST_EMU)
    atomicMax(p, v);
#else
    if (v > *p) *p = v;
#endif
}
SPX_HD void key_or(unsigned *p, unsigned v) {
#if defined(__CUDA_ARCH__) && !defined(SPX_HOST_EMU)
    atomicOr(p, v);
#else
    *p |= v;
#endif
}

// Geometry of one tile column as seen by one lane: pointer to cutout pixel
// (0, x) and whether that column exists in the source image.  Pixels outside
// the source image read as 0 (Cutout 'fill' mode, cutout.py:737-752).
struct TileCol {
    const float *p;        // address of cutout pixel (y=0, x) - may be out of range
    int pitch;
    bool colok;
};
SPX_HD TileCol tile_col(const PairState &ps, int s, int x, bool active) {
    TileCol t;
    t.pitch = int(ps.pitch[s]);
    t.colok = active && x >= ps.xlo && x < ps.xhi;
    t.p = ps.base[s] + ((long long)ps.y0[s] * t.pitch + (x + ps.x0[s]));
    return t;
}

// --------------------------------------------------------------------------
// statistics pre-pass (cc.py:131-156), one thread's share
// --------------------------------------------------------------------------
struct StatPartial {
    double cnt, sum, sq, rcnt, rsum, rsq;
    int flags;
};
SPX_HD StatPartial stat_zero() {
    StatPartial s; s.cnt = s.sum = s.sq = s.rcnt = s.rsum = s.rsq = 0.0; s.flags = 0; return s;
}
SPX_HD void stat_merge(StatPartial &a, const StatPartial &b) {
    a.cnt += b.cnt; a.sum += b.sum; a.sq += b.sq;
    a.rcnt += b.rcnt; a.rsum += b.rsum; a.rsq += b.rsq; a.flags |= b.flags;
}

// max(acc, |v|) that PROPAGATES NaN (one FMNMX.NAN with an |.| operand): the running
// "largest magnitude" of a batch of pixels is NaN / Inf iff one of them is
SPX_HD float absmax_nan(float acc, float v) {
#if defined(__CUDA_ARCH__) && !defined(SPX_HOST_EMU)
    float r;
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(acc), "f"(fabsf(v)));
    return r;
#else
    const float a = fabsf(v);
    return (acc != acc || a != a) ? NAN : (a > acc ? a : acc);
#endif
}

// non-finite pixels become 0; they raise the tile's status bit unless the
// caller asked for masked-pixel zeroing (align.py:661)
// internal status bit: a non-finite pixel was seen (reported or silently zeroed);
// never leaves the kernel
constexpr int kSawNonFinite = 0x40000000;

SPX_HD float sanitize(float v, int s, int report, int &flags) {
    if ((f32_bits(v) & 0x7f800000u) == 0x7f800000u) {
        flags |= kSawNonFinite | (report & (SPX_S_NONFINITE_REF << s));
        v = 0.f;
    }
    return v;
}

template <class C>
SPX_HD StatPartial stats_thread(const Ctx<C> &c, int tid) {
    const PairState &ps = *c.ps;
    StatPartial st = stat_zero();
    const int warp = warp_of(tid), lane = tid & 31;
    const int report = ps.report, ylo = ps.ylo, yhi = ps.yhi;
    int xlo = ps.xlo, xhi = ps.xhi;
    if (C::CL > 1) {                        // my share of the columns
        if (xlo < c.col_lo) xlo = c.col_lo;
        if (xhi > c.col_lo + C::NXL) xhi = c.col_lo + C::NXL;
    }
    const int pitch = int(ps.pitch[0]);
    const long long step = (long long)C::NW * pitch;
    int cnt = 0, rcnt = 0;
    // Pixels outside [ylo,yhi) x [xlo,xhi) are fill pixels: zero in every tile, so
    // they take part in no sum (cc.py:134-137 masks zeros).  The five tiles share
    // the box geometry, so one element offset serves all five base pointers.
#ifndef SPX_STATS_U
#define SPX_STATS_U 4
#endif
    constexpr int U = SPX_STATS_U;                           // rows in flight per lane
    for (int x = xlo + lane; x < xhi; x += 32) {
        const long long off = (long long)(ps.y0[0] + ylo + warp) * pitch + (x + ps.x0[0]);
        const float *p[5];
#pragma unroll
        for (int s = 0; s < 5; ++s) p[s] = ps.base[s] + off;
        for (int y = ylo + warp; y < yhi; y += U * C::NW) {
            float v[U][5];
#pragma unroll
            for (int u = 0; u < U; ++u) {                    // all loads first
                const bool ok = y + u * C::NW < yhi;
#pragma unroll
                for (int s = 0; s < 5; ++s) v[u][s] = ok ? ld_tile<C>(p[s] + u * step) : 0.f;
            }
#pragma unroll
            for (int s = 0; s < 5; ++s) p[s] += U * step;
            float wacc = 0.f;
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int s = 0; s < 5; ++s) wacc = absmax_nan(wacc, v[u][s]);
            const unsigned worst = f32_bits(wacc) & 0x7fffffffu;
            if (worst >= 0x7f800000u) {                      // rare: NaN / Inf present
#pragma unroll
                for (int u = 0; u < U; ++u)
#pragma unroll
                    for (int s = 0; s < 5; ++s) v[u][s] = sanitize(v[u][s], s, report, st.flags);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                bool any = false;
#pragma unroll
                for (int s = 1; s < 5; ++s) {
                    const bool nz = v[u][s] != 0.f;           // cc.py:134-137
                    const double d = double(v[u][s]);
                    cnt += nz ? 1 : 0; st.sum += d; st.sq += d * d;
                    any |= nz;
                }
                const double d = double(any ? v[u][0] : 0.f); // cc.py:150-154
                rcnt += any ? 1 : 0; st.rsum += d; st.rsq += d * d;
            }
        }
    }
    st.cnt = double(cnt); st.rcnt = double(rcnt);
    return st;
}

SPX_HD void stats_finalize(PairState &ps, const StatPartial &t) {
    ps.status |= t.flags;
    const bool zero = ps.cc_type == SPX_ZNCC;
    double mu = 0.0, sig = NAN, rmu = 0.0, rsig = NAN;
    if (t.cnt > 0.0) {
        mu = t.sum / t.cnt;
        double var = t.sq / t.cnt - mu * mu;
        sig = sqrt(var > 0.0 ? var : 0.0);
    }
    if (t.rcnt > 0.0) {
        rmu = t.rsum / t.rcnt;
        double var = t.rsq / t.rcnt - rmu * rmu;
        rsig = sqrt(var > 0.0 ? var : 0.0);
    }
    const float sf = float(sig), rf = float(rsig);
    const double sc = 1.0 / (double(sf) * double(rf));
    ps.mu = zero ? float(mu) : 0.f;
    ps.mu_ref = zero ? float(rmu) : 0.f;
    if (!(sc == sc) || !(fabs(sc) < 1e300) || sf == 0.f || rf == 0.f) {
        ps.status |= SPX_S_BAD_NORM;
        ps.rscale = 1.f;
    } else {
        ps.rscale = float(sc);
    }
    if (t.flags & ~kSawNonFinite) ps.status |= SPX_S_BAD_NORM;   // NaN/Inf poisons the statistics
}

// --------------------------------------------------------------------------
// column passes
// --------------------------------------------------------------------------
// C1: tiles (sa -> real part, sb -> imaginary part, sb < 0: none) placed at
// row offset yoff; E1-point FFT over n1 for every (column i, n2).
//
// FAST = full-size (ny = nx = NMAX) cutout lying entirely inside the source
// image.  The image tiles are then placed at offset NMAX/2 (a multiple of E2)
// instead of (n-1)/2, which makes the zero-padding pattern a compile-time
// property of n1: reference rows exist for n1 < NMAX/E2, image rows for
// n1 >= NMAX/(2*E2).  No bounds tests, no predicates, minimal address math.
// CHECK = false: the statistics pass has already scanned these tiles and found
// them finite, so the NaN/Inf tracking is skipped.
//
// lazy (statistics-free NCC path, see lazy_decide): 1 = also track the largest |pixel| of
// the pass; 2 = and look for evidence that the standard deviations of cc.py:139-154 are
// positive - two non-zero pixels of image00 / image10 with different values, resp. (in the
// reference pass, where four rows of image00 ride along as the mask) two reference pixels
// under the mask "image00 non-zero" with different values.  Two candidate pairs of rows
// per item are looked at; ANY hit in the CTA is enough.
SPX_HD bool lazy_evidence(float u, float w) {
    const float m = fmaxf(fabsf(u), fabsf(w));
    return fabsf(u - w) >= 1e-3f * m && m >= 1e-12f;
}
// CHUNK: only the items of 32-column chunk `chunk`, written to S as [PL][CP] (column = lane)
template <class C, bool FAST, bool IS_REF, bool CHECK = true, bool CHUNK = false>
SPX_HD int c1_pass(const Ctx<C> &c, int tid, int sa, int sb, int yoff, int col0,
                   int lazy = 0, int chunk = 0) {
    const PairState &ps = *c.ps;
    const int warp = warp_of(tid), lane = tid & 31;
    const int nx = FAST ? C::NXL : local_cols(c, ps.nx), ny = FAST ? C::NMAX : ps.ny;
    const int report = ps.report, ylo = ps.ylo, yhi = ps.yhi;
    const int nchunk = (nx + 31) >> 5;            // nx, i: LOCAL columns from here on
    const bool pair = !IS_REF;
    const bool zncc = ps.cc_type == SPX_ZNCC;
    const float mu = ps.mu, mu_ref = ps.mu_ref;
    constexpr int N1_LO = (FAST && !IS_REF) ? C::NMAX / (2 * C::E2) : 0;
    constexpr int N1_HI = (FAST && IS_REF) ? C::NMAX / C::E2 : C::E1;
    static_assert(!FAST || (C::NMAX / 2) % C::E2 == 0, "fast path needs E2 | NMAX/2");
    // candidate rows of the evidence search: two pairs around the middle of the rows an
    // item loads (the middle rows are the ones that exist on the generic path too)
    constexpr int MID = (N1_LO + N1_HI) / 2;
    constexpr int A0 = MID - 1, A1 = MID, B0 = (MID + 1 < N1_HI - 2) ? MID + 1 : N1_HI - 2, B1 = B0 + 1;
    static_assert(A0 >= N1_LO && B1 < N1_HI && B0 >= N1_LO, "evidence rows inside the item");
    const bool detect = lazy == 2;
    const WBases<C> wr = w_bases<C>(c);
    int flags = 0;
    unsigned amax = 0u;                            // bits of the largest finite |pixel|
    bool ev = false;
    // an item = (32-column chunk, n2): its loads, and everything after them
    auto load_item = [&](int wi, float (&va)[C::E1], float (&vb)[C::E1]) {
        const int n2 = wi % C::E2, i = (wi / C::E2) * 32 + lane;
        const bool active = FAST || i < nx;
        // reference pass with `detect`: four rows of image00 ride along in vb as the mask
        const bool second = pair || detect;
        const TileCol ta = tile_col(ps, sa, c.col_lo + i, active);
        const TileCol tb = tile_col(ps, pair ? sb : (detect ? 1 : sa), c.col_lo + i, active && second);
        const int ybase = n2 - yoff;
        const int rs = C::E2 * ta.pitch;                     // elements between loads
        const float *pa0 = ta.p + ybase * ta.pitch;
        const float *pb0 = tb.p + ybase * tb.pitch;
#if defined(__CUDA_ARCH__) && !defined(SPX_HOST_EMU) && !defined(SPX_NO_OPAQUE_PTR)
        // keep the two column pointers as materialised 64-bit byte addresses: each load is then
        // pointer + warp-uniform offset (two integer instructions) instead of the
        // (index + offset) * 4 + base the compiler re-associates to (four per load)
        asm volatile("" : "+l"(pa0), "+l"(pb0));
#endif
        static_for<0, C::E1>([&](auto N1) {
            constexpr int n1 = decltype(N1)::value;
            // generic path: the rows that exist depend on the item, so every row is a candidate
            [[maybe_unused]] constexpr bool cand = !FAST || n1 == A0 || n1 == A1 || n1 == B0 || n1 == B1;
            if constexpr (FAST) {
                if constexpr (n1 >= N1_LO && n1 < N1_HI) {
                    va[n1] = ld_tile_last<C>(pa0 + n1 * rs);
                    if constexpr (IS_REF) vb[n1] = (cand && detect) ? ld_tile<C>(pb0 + n1 * rs) : 0.f;
                    else vb[n1] = ld_tile_last<C>(pb0 + n1 * rs);
                } else {
                    va[n1] = 0.f; vb[n1] = 0.f;
                }
            } else {
                const int y = C::E2 * n1 + ybase;
                const bool ok = y >= ylo && y < yhi;         // warp-uniform
                va[n1] = (ok && ta.colok) ? ld_tile_last<C>(pa0 + n1 * rs) : 0.f;
                if constexpr (IS_REF) vb[n1] = (ok && tb.colok) ? ld_tile<C>(pb0 + n1 * rs) : 0.f;
                else vb[n1] = (ok && tb.colok) ? ld_tile_last<C>(pb0 + n1 * rs) : 0.f;
            }
        });
    };
    auto finish_item = [&](int wi, float (&va)[C::E1], float (&vb)[C::E1]) {
        const int n2 = wi % C::E2, i = (wi / C::E2) * 32 + lane;
        const bool active = FAST || i < nx;
        const int ybase = n2 - yoff;
        // one cheap test for "any NaN/Inf among my pixels": NaN-propagating max of |pixel|
        unsigned worst = 0u;
        if constexpr (CHECK) {
            float wacc = 0.f;
            static_for<N1_LO, N1_HI>([&](auto N1) {
                constexpr int n1 = decltype(N1)::value;
                wacc = absmax_nan(wacc, va[n1]);
                if (pair) wacc = absmax_nan(wacc, vb[n1]);
            });
            worst = f32_bits(wacc) & 0x7fffffffu;
        }
        if (CHECK && worst >= 0x7f800000u) {                 // rare
            float wacc = 0.f;
            static_for<0, C::E1>([&](auto N1) {
                constexpr int n1 = decltype(N1)::value;
                va[n1] = sanitize(va[n1], sa, report, flags);
                if (pair) vb[n1] = sanitize(vb[n1], sb, report, flags);
                wacc = absmax_nan(wacc, va[n1]);
                if (pair) wacc = absmax_nan(wacc, vb[n1]);
            });
            worst = f32_bits(wacc) & 0x7fffffffu;
        }
        if constexpr (CHECK) amax = amax > worst ? amax : worst;
        if (!FAST && detect) {
            // generic path: smallest / largest candidate value of this thread's column
            float tmin = INFINITY, tmax = -INFINITY;
            static_for<0, C::E1>([&](auto N1) {
                constexpr int n1 = decltype(N1)::value;
                const float a = va[n1], b = vb[n1];
                if constexpr (IS_REF) {
                    const bool m = b != 0.f && finitef(b);
                    tmin = fminf(tmin, m ? a : INFINITY); tmax = fmaxf(tmax, m ? a : -INFINITY);
                    vb[n1] = 0.f;
                } else {
                    const bool za = a != 0.f, zb = b != 0.f;
                    tmin = fminf(tmin, fminf(za ? a : INFINITY, zb ? b : INFINITY));
                    tmax = fmaxf(tmax, fmaxf(za ? a : -INFINITY, zb ? b : -INFINITY));
                }
            });
            ev |= tmax >= tmin && lazy_evidence(tmin, tmax);
        }
        if (FAST && detect) {
            if constexpr (IS_REF) {
                // reference pixels where image00 is non-zero (cc.py:150-154; a non-finite
                // image00 pixel is either reported or zeroed - not part of the mask)
                const bool ma = vb[A0] != 0.f && finitef(vb[A0]) && vb[A1] != 0.f && finitef(vb[A1]);
                const bool mb = vb[B0] != 0.f && finitef(vb[B0]) && vb[B1] != 0.f && finitef(vb[B1]);
                ev |= (ma && lazy_evidence(va[A0], va[A1])) || (mb && lazy_evidence(va[B0], va[B1]));
                vb[A0] = 0.f; vb[A1] = 0.f; vb[B0] = 0.f; vb[B1] = 0.f;
            } else {
                // non-zero pixels of image00 / image10 (cc.py:134-137)
                ev |= (va[A0] != 0.f && va[A1] != 0.f && lazy_evidence(va[A0], va[A1])) ||
                      (vb[B0] != 0.f && vb[B1] != 0.f && lazy_evidence(vb[B0], vb[B1]));
            }
        }
        cf x[C::E1];
        if (zncc) {
            static_for<0, C::E1>([&](auto N1) {
                constexpr int n1 = decltype(N1)::value;
                const int y = C::E2 * n1 + ybase;
                float a = va[n1], b = vb[n1];
                if (IS_REF) { if (active && y >= 0 && y < ny) a -= mu_ref; }      // cc.py:152-153
                else {
                    if (a != 0.f) a -= mu;                                          // cc.py:145-146
                    if (b != 0.f) b -= mu;
                }
                x[n1] = cmake(a, b);
            });
        } else {
            static_for<0, C::E1>([&](auto N1) {
                constexpr int n1 = decltype(N1)::value;
                x[n1] = cmake(va[n1], vb[n1]);
            });
        }
        RegFFT<C::E1, false>::run(x);
        if (CHUNK) {                                         // (idle lanes write zeros)
            static_for<0, C::E1>([&](auto K1) {
                constexpr int k1 = decltype(K1)::value;
                c.S[(k1 * C::E2 + n2) * C::CP + lane] = cmul(x[k1], tw2<C>(n2, k1));
            });
        } else if (active) {
            const int gi = c.col_lo + i + col0;              // column inside W
            static_for<0, C::E1>([&](auto K1) {
                constexpr int k1 = decltype(K1)::value;
                cf *wb = wr.p[k1 % C::CL];                       // owner of the k1 group
                wb[((k1 / C::CL) * C::E2 + n2) * C::WP + gi] = cmul(x[k1], tw2<C>(n2, k1));
            });
        }
    };
    const int nitems = CHUNK ? (chunk + 1) * C::E2 : nchunk * C::E2;
    const int item0 = CHUNK ? chunk * C::E2 : 0;
    // 16 loads per item on the fast path: two items fit the 64 px class's registers
    // (+1.6 %; the 32 and 128 px classes spill and lose 2-4 %)
#if !defined(SPX_HOST_EMU) && !defined(SPX_NO_C1_PAIR)
    constexpr bool kPairItems = FAST && C::NMAX == 64 && !C::STAGE;   // (staged tiles: no L2 latency to hide)
#else
    constexpr bool kPairItems = false;
#endif
    if constexpr (kPairItems) {
        // two items per iteration: the tile loads of the second (L2 latency) are in flight
        // while the first goes through its FFT
        for (int wi = item0 + warp; wi < nitems; wi += 2 * C::NW) {
            float va0[C::E1], vb0[C::E1], va1[C::E1], vb1[C::E1];
            const bool two = wi + C::NW < nitems;            // warp-uniform
            load_item(wi, va0, vb0);
            if (two) load_item(wi + C::NW, va1, vb1);
            finish_item(wi, va0, vb0);
            if (two) finish_item(wi + C::NW, va1, vb1);
        }
    } else {
        SPX_UNROLL_BIG
        for (int wi = item0 + warp; wi < nitems; wi += C::NW) {
            float va[C::E1], vb[C::E1];
            load_item(wi, va, vb);                           // all loads first (independent)
            finish_item(wi, va, vb);                         // arithmetic afterwards
        }
    }
    if (lazy) {
        unsigned evbit = ev ? (IS_REF ? 2u : 1u) : 0u;
#if defined(__CUDA_ARCH__) && !defined(SPX_HOST_EMU)
        evbit = __reduce_or_sync(0xffffffffu, evbit);
        amax = __reduce_max_sync(0xffffffffu, amax);
        if (lane == 0)
#endif
        {
            // det[0]: evidence bits, det[1]: largest |pixel| seen; cluster mode: both copies
            for (int r = 0; r < C::CL; ++r) {
                Xchg *x = X_of<C>(c, r);
                if (evbit) key_or(&x->det[c.dslot][0], evbit);
                key_max(&x->det[c.dslot][1], amax);
            }
        }
    }
    return flags;
}

// Statistics-free NCC (cc_type NCC, no surface requested).  The pooled sigma and
// sigma_ref of cc.py:139-154 scale the whole surface by ONE positive number, and the
// result of the path - argmax, the quadratic-fit position (homogeneous of degree 0 in
// the samples) and every sign test of centroid.py:216-236 - does not depend on a
// positive scale.  The statistics only decide the degenerate cases (no non-zero pixel
// -> std of an empty set -> NaN surface; sigma or sigma_ref == 0 -> division by zero),
// so it is enough to PROVE both deviations positive.  Evidence: two pixels of the set a
// deviation is taken over (non-zero pixels of the four phases; reference pixels under
// the union mask) whose values differ by >= 1e-3 of the larger magnitude.  If the value
// M of largest magnitude in the set and a value v0 with |v0 - M| >= ~1e-3 |M| both
// exist - which two such pixels imply - then var >= (1e-3 M)^2 / (2 n) while the mean
// square is <= M^2, a ratio >= 5e-7 / n that the float64 sums of the exact path
// (relative error ~ 1e-16) cannot cancel for any n up to 4 * 256^2.  With every |pixel| in
// [0, 1e12] (and the evidence pixels >= 1e-12) neither the unnormalised surface nor the
// exact path's scale leaves the float32 range.  Anything else - all-zero or constant
// tiles, a single unmasked pixel, extreme magnitudes - falls back to the exact path.
SPX_HD bool lazy_decide(const unsigned (&det)[2], bool need_evidence) {
#ifdef SPX_LAZY_NODETECT
    need_evidence = false;
#endif
#ifdef SPX_LAZY_NOREF
    return (!need_evidence || (det[0] & 1u) == 1u) && det[1] <= 0x5368d4a5u;
#endif
    return (!need_evidence || (det[0] & 3u) == 3u) && det[1] <= 0x5368d4a5u /* 1e12f */;
}

// C2: E2-point FFT over n2 for every (column, k1 group of mine), in place.
// (FAST: full-size tile - nx, ny and the surface width are compile-time constants)
// (ch_lo, ch_hi: only the 32-column chunks [ch_lo, ch_hi) - split-phase cluster barriers)
// DIRECT_W (slab classes, CHUNK): the outputs go straight to the W slab instead of back into
// the chunk - lanes along x, 256 contiguous bytes per store instruction - which saves the
// copy-out pass and its barrier; the reference only stores the rows with ky <= P/2.
template <class C, bool FAST = false, bool CHUNK = false, bool DIRECT_W = false, bool IS_REF = false>
SPX_HD void c2_pass(const Ctx<C> &c, int tid, int col0, int chunk = 0, int ch_lo = 0, int ch_hi = 1 << 20) {
    const int warp = warp_of(tid), lane = tid & 31;
    const int nx = FAST ? C::NMAX : c.ps->nx;
    const int nchunk = (nx + 31) >> 5;
    constexpr int WPc = CHUNK ? C::CP : C::WP;
    cf *wbase = CHUNK ? c.S : c.W;
    if (ch_hi > nchunk) ch_hi = nchunk;
    const int item0 = CHUNK ? chunk * C::E1L : ch_lo * C::E1L;
    const int item1 = CHUNK ? (chunk + 1) * C::E1L : ch_hi * C::E1L;
    SPX_UNROLL_SMALL
    for (int wi = item0 + warp; wi < item1; wi += C::NW) {
        const int j = wi % C::E1L, i = (wi / C::E1L) * 32 + lane;
        if (i >= nx) continue;
        const int col = CHUNK ? lane : i + col0;
        cf x[C::E2];
        static_for<0, C::E2>([&](auto N2) {
            constexpr int n2 = decltype(N2)::value;
            x[n2] = wbase[(j * C::E2 + n2) * WPc + col];
        });
        RegFFT<C::E2, false>::run(x);
        if constexpr (CHUNK && DIRECT_W) {
            cf *wout = c.W + (j * C::E2) * C::WP + i + col0;
            static_for<0, C::E2>([&](auto K2) {
                constexpr int k2 = decltype(K2)::value;
                if (!IS_REF || j + C::E1 * k2 <= C::P / 2) wout[k2 * C::WP] = x[k2];
            });
        } else {
            static_for<0, C::E2>([&](auto K2) {
                constexpr int k2 = decltype(K2)::value;
                wbase[(j * C::E2 + k2) * WPc + col] = x[k2];
            });
        }
    }
}

// IC1: E2-point inverse FFT over k2 (+ inverse twiddle), in place.
// DIRECT_W (slab classes, CHUNK): the inputs come straight from the W slab (no copy-in pass),
// the outputs go to the chunk in S for IC2.
template <class C, bool FAST = false, bool CHUNK = false, bool DIRECT_W = false>
SPX_HD void ic1_pass(const Ctx<C> &c, int tid, int col0, int chunk = 0, int ch_lo = 0, int ch_hi = 1 << 20) {
    const int warp = warp_of(tid), lane = tid & 31;
    const int nx = FAST ? C::NMAX : c.ps->nx;
    const int nchunk = (nx + 31) >> 5;
    constexpr int WPc = CHUNK ? C::CP : C::WP;
    cf *wbase = CHUNK ? c.S : c.W;
    if (ch_hi > nchunk) ch_hi = nchunk;
    const int item0 = CHUNK ? chunk * C::E1L : ch_lo * C::E1L;
    const int item1 = CHUNK ? (chunk + 1) * C::E1L : ch_hi * C::E1L;
    SPX_UNROLL_SMALL
    for (int wi = item0 + warp; wi < item1; wi += C::NW) {
        const int j = wi % C::E1L, l = (wi / C::E1L) * 32 + lane;
        const int k1 = j * C::CL + c.rank;
        if (l >= nx) continue;
        const int col = CHUNK ? lane : l + col0;
        cf x[C::E2];
        if constexpr (CHUNK && DIRECT_W) {
            const cf *win = c.W + (j * C::E2) * C::WP + l + col0;
            static_for<0, C::E2>([&](auto K2) {
                constexpr int k2 = decltype(K2)::value;
                x[k2] = win[k2 * C::WP];
            });
        } else {
            static_for<0, C::E2>([&](auto K2) {
                constexpr int k2 = decltype(K2)::value;
                x[k2] = wbase[(j * C::E2 + k2) * WPc + col];
            });
        }
        RegFFT<C::E2, true>::run(x);
        static_for<0, C::E2>([&](auto N2) {
            constexpr int n2 = decltype(N2)::value;
            wbase[(j * C::E2 + n2) * WPc + col] = cmul(x[n2], tw2<C, true>(n2, k1));
        });
    }
}

// Store two surface samples.  `scratch`: the destination is the per-CTA scratch slab that
// only the fit reads back (nobody asked for the surface): the store carries an L2
// evict-last policy so that the streaming tile loads do not push the slab out to HBM.
SPX_HD unsigned long long l2_evict_last_policy() {
#if defined(__CUDA_ARCH__) && !defined(SPX_HOST_EMU)
    unsigned long long pol;
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
#else
    return 0ull;
#endif
}
SPX_HD void st_surface(float *p, float v0, float v1, bool scratch, unsigned long long pol) {
#if defined(__CUDA_ARCH__) && !defined(SPX_HOST_EMU)
    if (scratch) {
        asm volatile("st.global.L2::cache_hint.v2.f32 [%0], {%1, %2}, %3;"
                     :: "l"(p), "f"(v0), "f"(v1), "l"(pol) : "memory");
        return;
    }
#endif
    (void)scratch; (void)pol;
    *reinterpret_cast<cf *>(p) = cmake(v0, v1);
}

// IC2: E1-point inverse FFT over k1; outputs n = E2*n1+n2 < ny are the lags
// n - yc of phases (a=0, b) [real part] and (a=1, b) [imaginary part].
// Inputs were sanitised, so the surface is finite and a plain ">" scan in
// increasing index order implements numpy's first-maximum rule per item.
template <class C, bool FAST = false, bool CHUNK = false>
SPX_HD Best ic2_pass(const Ctx<C> &c, int tid, int b, int shift, int col0, int chunk = 0) {
    const PairState &ps = *c.ps;
    const int warp = warp_of(tid), lane = tid & 31;
    const int ny = FAST ? C::NMAX : ps.ny, nx = FAST ? C::NXL : local_cols(c, ps.nx);
    const int nchunk = (nx + 31) >> 5;
    const int w2 = FAST ? 2 * C::NMAX : 2 * ps.nx;
    float *icc = c.icc;
    const bool keep = ps.keep != 0;              // nobody asked for the surface
    // (the 80-register classes have no room for the 64-bit policy across the item loop)
    const unsigned long long pol = C::T * C::MINB <= 512 ? l2_evict_last_policy() : 0ull;
    const int cbase = 2 * c.col_lo;              // my columns start at surface column 2*col_lo
    const WBases<C> wr = w_bases<C>(c);
    Best best = best_init();
    const int item0 = CHUNK ? chunk * C::E2 : 0, item1 = CHUNK ? (chunk + 1) * C::E2 : nchunk * C::E2;
    SPX_UNROLL_BIG
    for (int wi = item0 + warp; wi < item1; wi += C::NW) {
        const int n2 = wi % C::E2, l = (wi / C::E2) * 32 + lane;
        if (l >= nx) continue;
        cf x[C::E1];
        const int gl = c.col_lo + l + col0;                 // column inside W
        static_for<0, C::E1>([&](auto K1) {
            constexpr int k1 = decltype(K1)::value;
            if constexpr (CHUNK) {
                x[k1] = c.S[(k1 * C::E2 + n2) * C::CP + lane];
            } else {
                const cf *wb = wr.p[k1 % C::CL];
                x[k1] = wb[((k1 / C::CL) * C::E2 + n2) * C::WP + gl];
            }
        });
        RegFFT<C::E1, true>::run(x);
        // sample n = E2*n1 + n2 is lag (n - shift) - yc -> surface row 2*(n-shift)+b
        const int nb = n2 - shift;
        const int idx0 = (2 * nb + b) * w2 + 2 * l + cbase;
        float bv = -INFINITY;
        int bi = 0x7fffffff;
        static_for<0, C::E1>([&](auto N1) {
            constexpr int n1 = decltype(N1)::value;
            if (unsigned(C::E2 * n1 + nb) < unsigned(ny)) {
                const float v0 = x[n1].x, v1 = x[n1].y;
                const int idx = idx0 + n1 * (2 * C::E2) * w2;
                if (SurfaceInW<C>::value && keep)       // in place: the slot this thread read
                    c.W[(n1 * C::E2 + n2) * C::WP + gl] = cmake(v0, v1);
                else
                    st_surface(icc + idx, v0, v1, keep && C::T * C::MINB <= 512, pol);
                if (v0 > bv) { bv = v0; bi = idx; }
                if (v1 > bv) { bv = v1; bi = idx + 1; }
            }
        });
        Best cand; cand.v = bv; cand.idx = bi;
        if (best_better(cand, best)) best = cand;
    }
    return best;
}

// --------------------------------------------------------------------------
// row passes; lane <-> row.  A warp-item is (chunk of 32 rows, n2 or k1); the
// passes loop over the chunks [cb, ce).  INPLACE: the row's P samples live in
// W itself (column = padded position), all chunks in one go.  Otherwise the
// exchange goes through the scratch S and the caller runs one chunk at a time.
// --------------------------------------------------------------------------
template <class C, bool IS_REF>
SPX_HD int row_of(const Ctx<C> &c, int chunk, int lane) {
    const int k = chunk * 32 + lane;
    if (IS_REF) return k < c.n_rrow ? int(c.rrow[k]) : -1;
#ifndef SPX_NO_ROW_PAIRING
    if constexpr (C::FUSE && C::E2 == 16 && C::E1 % 2 == 0) {
        // slab classes: a batch is the k1 group g together with group E1 - g (0 with E1/2):
        // row (k1, k2) holds ky = k1 + E1*k2 and row (E1 - k1, E2 - 1 - k2) holds P - ky, so
        // the two rows that read the same R row (direct and mirrored) sit in the SAME batch
        // and R comes from HBM once per phase pair instead of twice
        const int g = lane < 16 ? (chunk == 0 ? 0 : chunk) : (chunk == 0 ? C::E1 / 2 : C::E1 - chunk);
        return g * 16 + (lane & 15);
    }
#endif
    return k < C::PL ? k : -1;
}
// row of the (local part of the) stored half spectrum that holds frequency row q
template <class C> SPX_HD int rmap(int q) { return q / C::CL; }
template <class C>
SPX_HD cf *srow_of(const Ctx<C> &c, int r, int lane) {
    return C::INPLACE ? c.W + (r >= 0 ? r : 0) * C::WP : c.S + lane * C::SP;
}

// Staged mode (W in global memory): coalesced copy of one batch of rows between
// W (columns = input / lag index) and the scratch S (columns = padded position);
// one warp per row, lanes along the row.
// slab classes with a line pitch: the global side of the copies moves two values per access
template <class C> struct Wide128 {
#ifndef SPX_NO_WIDE_COPIES
    static constexpr bool value = !C::W_SMEM && !C::INPLACE && C::WP % 2 == 0 && C::CL == 1;
#else
    static constexpr bool value = false;
#endif
};
template <class C, bool IS_REF>
SPX_HD void stage_in(const Ctx<C> &c, int tid, int chunk, int xoff) {
    const int warp = warp_of(tid), lane = tid & 31;
    const int nx = c.ps->nx;
    if constexpr (Wide128<C>::value) {
        constexpr int NP = ((C::NMAX + 1) / 2 + 31) / 32;    // value PAIRS of a row per lane
        // RB rows per iteration: every 128-bit load of all of them is issued before the first
        // store - RB * NP independent loads in flight (four rows: +1 % at 170 px, -1 % at 256 px)
#ifndef SPX_STAGE_ROWS
        constexpr int RB = 2;
#else
        constexpr int RB = SPX_STAGE_ROWS;
#endif
        for (int k = warp; k < 32; k += RB * C::NW) {
            int r[RB];
            cf2 v[RB][NP];
            static_for<0, RB>([&](auto B) {
                constexpr int b = decltype(B)::value;
                r[b] = k + b * C::NW < 32 ? row_of<C, IS_REF>(c, chunk, k + b * C::NW) : -1;
                const cf2 *src = reinterpret_cast<const cf2 *>(c.W + (r[b] >= 0 ? r[b] : 0) * C::WP);
                static_for<0, NP>([&](auto Q) {
                    constexpr int q = decltype(Q)::value;
                    const int i = lane + 32 * q;
                    cf2 z; z.a = z.b = cmake(0.f, 0.f);
                    v[b][q] = (r[b] >= 0 && 2 * i < nx) ? src[i] : z;
                });
            });
            static_for<0, RB>([&](auto B) {
                constexpr int b = decltype(B)::value;
                cf *d = c.S + (k + b * C::NW) * C::SP + xoff;
                static_for<0, NP>([&](auto Q) {
                    constexpr int q = decltype(Q)::value;
                    const int i = 2 * (lane + 32 * q);
                    if (r[b] >= 0 && i < nx) { d[i] = v[b][q].a; if (i + 1 < nx) d[i + 1] = v[b][q].b; }
                });
            });
        }
        return;
    }
    constexpr int NI = (C::NMAX + 31) / 32;      // elements of a row per lane
    // two rows per iteration, every load of both issued before the first store (the slab
    // is L2 / HBM resident: 2 * NI independent loads in flight per thread)
    for (int k = warp; k < 32; k += 2 * C::NW) {
        const int r0 = row_of<C, IS_REF>(c, chunk, k);
        const int r1 = k + C::NW < 32 ? row_of<C, IS_REF>(c, chunk, k + C::NW) : -1;
        const cf *s0 = c.W + (r0 >= 0 ? r0 : 0) * C::WP, *s1 = c.W + (r1 >= 0 ? r1 : 0) * C::WP;
        cf v0[NI], v1[NI];
        static_for<0, NI>([&](auto Q) {
            constexpr int q = decltype(Q)::value;
            const int i = lane + 32 * q;
            v0[q] = (r0 >= 0 && i < nx) ? s0[i] : cmake(0.f, 0.f);
            v1[q] = (r1 >= 0 && i < nx) ? s1[i] : cmake(0.f, 0.f);
        });
        cf *d0 = c.S + k * C::SP + xoff, *d1 = c.S + (k + C::NW) * C::SP + xoff;
        static_for<0, NI>([&](auto Q) {
            constexpr int q = decltype(Q)::value;
            const int i = lane + 32 * q;
            if (r0 >= 0 && i < nx) d0[i] = v0[q];
            if (r1 >= 0 && i < nx) d1[i] = v1[q];
        });
    }
}
template <class C>
SPX_HD void stage_out(const Ctx<C> &c, int tid, int chunk, int shift) {
    const int warp = warp_of(tid), lane = tid & 31;
    const int nx = c.ps->nx;
    for (int k = warp; k < 32; k += C::NW) {
        const int r = row_of<C, false>(c, chunk, k);
        if (r < 0) continue;
        const cf *src = c.S + k * C::SP + shift;
        cf *dst = c.W + r * C::WP;
        if constexpr (Wide128<C>::value) {
            // (an odd nx writes one value past the row's last lag: inside the pitch, never read)
            for (int l = 2 * lane; l < nx; l += 64) {
                cf2 v; v.a = src[l]; v.b = src[l + 1];
                *reinterpret_cast<cf2 *>(dst + l) = v;
            }
        } else {
            for (int l = lane; l < nx; l += 32) dst[l] = src[l];
        }
    }
}

// Fused column passes (C::FUSE): one 32-column chunk of W lives in S as [PL][CP]
// (column = lane) between C1 and C2 resp. IC1 and IC2; these copies move it to / from the
// slab - one warp per row, 256-byte row segments (chunk_in issues all loads of a thread
// before its stores).  The reference only needs the rows with ky <= P/2 (its row stage skips the rest).
template <class C, bool IS_REF>
SPX_HD void chunk_out(const Ctx<C> &c, int tid, int chunk, int col0) {
    const int warp = warp_of(tid), lane = tid & 31;
    if constexpr (Wide128<C>::value && C::PL % (2 * C::NW) == 0) {
        // half a warp per row, two values per lane: 512 contiguous bytes of S, 2 x 256 of W
        const int half = lane >> 4, col = chunk * 32 + 2 * (lane & 15);
        if (col >= c.ps->nx) return;
        constexpr int RW = C::PL / (2 * C::NW);
        static_for<0, RW>([&](auto K) {
            constexpr int k = decltype(K)::value;
            const int r = 2 * (warp + C::NW * k) + half;
            if (!IS_REF || ky_of_row<C>(r, 0) <= C::P / 2)
                *reinterpret_cast<cf2 *>(c.W + r * C::WP + col + col0) =
                    *reinterpret_cast<const cf2 *>(c.S + r * C::CP + 2 * (lane & 15));
        });
        return;
    }
    const int col = chunk * 32 + lane;
    if (col >= c.ps->nx) return;
    constexpr int RW = (C::PL + C::NW - 1) / C::NW;
    cf *dst = c.W + col + col0;
    static_for<0, RW>([&](auto K) {
        constexpr int k = decltype(K)::value;
        const int r = warp + C::NW * k;                       // warp-uniform
        if (r < C::PL && (!IS_REF || ky_of_row<C>(r, 0) <= C::P / 2))
            dst[r * C::WP] = c.S[r * C::CP + lane];
    });
}
template <class C>
SPX_HD void chunk_in(const Ctx<C> &c, int tid, int chunk, int col0) {
    const int warp = warp_of(tid), lane = tid & 31;
    if constexpr (Wide128<C>::value && C::PL % (2 * C::NW) == 0) {
        const int half = lane >> 4, col = chunk * 32 + 2 * (lane & 15);
        const bool ok = col < c.ps->nx;
        constexpr int RW = C::PL / (2 * C::NW);               // row pairs per warp
        constexpr int G = RW < 12 ? RW : 12;                  // 128-bit loads in flight per thread
        const cf *src = c.W + (ok ? col + col0 : 0);
        static_for<0, (RW + G - 1) / G>([&](auto GG) {
            constexpr int g = decltype(GG)::value;
            cf2 v[G];
            static_for<0, G>([&](auto K) {
                constexpr int k = g * G + decltype(K)::value;
                const int r = 2 * (warp + C::NW * k) + half;
                cf2 z; z.a = z.b = cmake(0.f, 0.f);
                v[k - g * G] = (ok && k < RW) ? *reinterpret_cast<const cf2 *>(src + r * C::WP) : z;
            });
            static_for<0, G>([&](auto K) {
                constexpr int k = g * G + decltype(K)::value;
                const int r = 2 * (warp + C::NW * k) + half;
                if (ok && k < RW)
                    *reinterpret_cast<cf2 *>(c.S + r * C::CP + 2 * (lane & 15)) = v[k - g * G];
            });
        });
        return;
    }
    const int col = chunk * 32 + lane;
    const bool ok = col < c.ps->nx;
    constexpr int RW = (C::PL + C::NW - 1) / C::NW;          // rows per warp
    constexpr int G = RW < 16 ? RW : 16;                     // loads in flight per thread
    const cf *src = c.W + (ok ? col + col0 : 0);
    static_for<0, (RW + G - 1) / G>([&](auto GG) {
        constexpr int g = decltype(GG)::value;
        cf v[G];
        static_for<0, G>([&](auto K) {
            constexpr int k = g * G + decltype(K)::value;
            const int r = warp + C::NW * k;
            v[k - g * G] = (ok && k < RW && r < C::PL) ? src[r * C::WP] : cmake(0.f, 0.f);
        });
        static_for<0, G>([&](auto K) {
            constexpr int k = g * G + decltype(K)::value;
            const int r = warp + C::NW * k;
            if (k < RW && r < C::PL) c.S[r * C::CP + lane] = v[k - g * G];
        });
    });
}

// L2 prefetches for the slab classes: the next row batch / column chunk of W (and the R rows
// its X2 pass will read) is pulled from HBM into L2 while the current one is transformed, so
// the copy passes pay the L2 latency instead of the DRAM latency.  -DSPX_NO_SLAB_PREFETCH=1
// switches them off (A/B knob).
SPX_HD void l2_prefetch(const void *p) {
#if defined(__CUDA_ARCH__) && !defined(SPX_HOST_EMU) && !defined(SPX_NO_SLAB_PREFETCH)
    asm volatile("prefetch.global.L2 [%0];" :: "l"(p));
#else
    (void)p;
#endif
}
// the 128-byte lines of [p, p + bytes), dealt to the lanes of a warp
SPX_HD void l2_prefetch_span(const void *p, int bytes, int lane) {
#if defined(__CUDA_ARCH__) && !defined(SPX_HOST_EMU)
    const unsigned long long lo = reinterpret_cast<unsigned long long>(p), hi = lo + bytes;
    for (unsigned long long q = (lo & ~127ull) + 128ull * lane; q < hi; q += 128ull * 32)
        l2_prefetch(reinterpret_cast<const void *>(q));
#else
    (void)p; (void)bytes; (void)lane;
#endif
}
template <class C, bool IS_REF>
SPX_HD void prefetch_row_batch(const Ctx<C> &c, int tid, int chunk) {
    if (chunk >= (IS_REF ? C::RCHUNK : C::NCHUNK)) return;
    const int warp = warp_of(tid), lane = tid & 31;
    const int nx = c.ps->nx;
    for (int k = warp; k < 32; k += C::NW) {
        const int r = row_of<C, IS_REF>(c, chunk, k);
        if (r < 0) continue;
        l2_prefetch_span(c.W + r * C::WP, nx * 8, lane);
        if (!IS_REF) {                                       // the R row of x2_pass
            const int ky = ky_of_row<C>(r, c.rank);
            const int q = ky > C::P / 2 ? C::P - ky : ky;
            l2_prefetch_span(c.R + rmap<C>(q) * C::RP, C::P * 8, lane);
        }
    }
}
template <class C>
SPX_HD void prefetch_col_chunk(const Ctx<C> &c, int tid, int chunk, int col0) {
    if (chunk * 32 >= c.ps->nx) return;
    for (int r = tid; r < C::PL; r += C::T) {
        const cf *p = c.W + r * C::WP + chunk * 32 + col0;   // 256 bytes: up to three lines
        l2_prefetch(p); l2_prefetch(p + 16); l2_prefetch(p + 31);
    }
}

// X1: E1-point FFT along x of the nx input samples (padded position m = E2*n1+n2,
// input column m - xoff)
template <class C, bool FAST, bool IS_REF>
SPX_HD void x1_pass(const Ctx<C> &c, int tid, int cb, int ce, int xoff) {
    const int warp = warp_of(tid), lane = tid & 31;
    const int nx = FAST ? C::NMAX : c.ps->nx;
    constexpr int N1_LO = (FAST && !IS_REF) ? C::NMAX / (2 * C::E2) : 0;
    constexpr int N1_HI = (FAST && IS_REF) ? C::NMAX / C::E2 : C::E1;
    SPX_UNROLL_BIG
    for (int wi = cb * C::E2 + warp; wi < ce * C::E2; wi += C::NW) {
        const int chunk = wi / C::E2, n2 = wi - chunk * C::E2;
        const int r = row_of<C, IS_REF>(c, chunk, lane);
        if (C::INPLACE && r < 0) continue;
        // the P-wide row buffer (the W row itself / a row of S) is indexed by padded
        // position; DIRECT mode reads the nx input samples straight from the W row
        cf *srow = srow_of<C>(c, r, lane);
        const cf *in = C::DIRECT ? c.W + (r >= 0 ? r : 0) * C::WP - xoff : srow;
        cf x[C::E1];
        static_for<0, C::E1>([&](auto N1) {
            constexpr int n1 = decltype(N1)::value;
            if constexpr (FAST) {
                if constexpr (n1 >= N1_LO && n1 < N1_HI)
                    x[n1] = (r < 0) ? cmake(0.f, 0.f) : in[C::E2 * n1 + n2];
                else
                    x[n1] = cmake(0.f, 0.f);
            } else {
                const int i = C::E2 * n1 + n2 - xoff;
                x[n1] = (r >= 0 && i >= 0 && i < nx) ? in[C::E2 * n1 + n2] : cmake(0.f, 0.f);
            }
        });
        RegFFT<C::E1, false>::run(x);
        static_for<0, C::E1>([&](auto K1) {
            constexpr int k1 = decltype(K1)::value;
            srow[k1 * C::E2 + n2] = cmul(x[k1], tw2<C>(n2, k1));
        });
    }
}

// X2 for the reference image: finish the forward transform and store
// conj(R) * scale for ky <= P/2.
template <class C>
SPX_HD void x2_ref_pass(const Ctx<C> &c, int tid, int cb, int ce) {
    const int warp = warp_of(tid), lane = tid & 31;
    // normalisation scale and the 1/P^2 of the unnormalised inverse FFTs
    const float sc = c.ps->rscale * (1.0f / (float(C::P) * float(C::P)));
    SPX_UNROLL_SMALL
    for (int wi = cb * C::E1 + warp; wi < ce * C::E1; wi += C::NW) {
        const int chunk = wi / C::E1, k1 = wi - chunk * C::E1;
        const int r = row_of<C, true>(c, chunk, lane);
        if (C::INPLACE && r < 0) continue;
        const cf *srow = srow_of<C>(c, r, lane);
        cf x[C::E2];
        static_for<0, C::E2>([&](auto N2) {
            constexpr int n2 = decltype(N2)::value;
            x[n2] = srow[k1 * C::E2 + n2];
        });
        RegFFT<C::E2, false>::run(x);
        if (r >= 0) {
            cf *Rrow = c.R + rmap<C>(ky_of_row<C>(r, c.rank)) * C::RP + k1 * C::E2;
            if constexpr (!C::R_SMEM && C::RP % 2 == 0 && C::E2 % 2 == 0) {
                static_for<0, C::E2 / 2>([&](auto J) {           // line pitch: 128-bit stores
                    constexpr int j = decltype(J)::value;
                    cf2 v;
                    v.a = cmake(x[2 * j].x * sc, -x[2 * j].y * sc);
                    v.b = cmake(x[2 * j + 1].x * sc, -x[2 * j + 1].y * sc);
                    reinterpret_cast<cf2 *>(Rrow)[j] = v;
                });
            } else
            static_for<0, C::E2>([&](auto K2) {
                constexpr int k2 = decltype(K2)::value;
                Rrow[k2] = cmake(x[k2].x * sc, -x[k2].y * sc);
            });
        }
    }
}

// X2: forward E2-FFT, multiply by conj(R[ky, kx]), inverse E2-FFT (+twiddle)
template <class C>
SPX_HD void x2_pass(const Ctx<C> &c, int tid, int cb, int ce) {
    const int warp = warp_of(tid), lane = tid & 31;
    const cf *Rmir = R_of<C>(c, (C::CL - c.rank) % C::CL);
    SPX_UNROLL_SMALL
    for (int wi = cb * C::E1 + warp; wi < ce * C::E1; wi += C::NW) {
        const int chunk = wi / C::E1, k1 = wi - chunk * C::E1;
        const int r = row_of<C, false>(c, chunk, lane);
        if (C::INPLACE && r < 0) continue;
        cf *srow = srow_of<C>(c, r, lane);
        const int ky = r >= 0 ? ky_of_row<C>(r, c.rank) : 0;
        const bool flip = ky > C::P / 2;
        // the mirrored row P - ky belongs to CTA (CL - rank) % CL (see CcCfg)
        const cf *Rrow = (flip ? Rmir : c.R) + rmap<C>(flip ? C::P - ky : ky) * C::RP;
        cf x[C::E2];
        static_for<0, C::E2>([&](auto N2) {
            constexpr int n2 = decltype(N2)::value;
            x[n2] = srow[k1 * C::E2 + n2];
        });
        RegFFT<C::E2, false>::run(x);
        // column of kx = k1 + E1*k2 is k1*E2 + k2; the mirrored frequency
        // (P - kx) % P sits at (E1-k1)*E2 + (E2-1-k2) for k1 > 0 and at
        // (E2-k2) % E2 for k1 == 0  (all warp-uniform)
        const int col_d = k1 * C::E2;
        const int col_f = k1 ? (C::E1 - k1) * C::E2 + C::E2 - 1 : C::E2;
        const cf *Rd = Rrow + (flip ? col_f : col_d);
#if defined(__CUDA_ARCH__) && !defined(SPX_HOST_EMU) && !defined(SPX_NO_R128)
        // R in a global slab (line pitch, so every k1 block of a row is 16-byte aligned): two
        // values per 128-bit load.  A mirrored row reads the block (E1-k1)*E2 .. +E2-1 backwards:
        // the pair at 14-2j holds the values for k2 = 2j+1 and 2j (k1 = 0 mixes element 0 in and
        // keeps the scalar loads).
        if (!C::R_SMEM && C::RP % 2 == 0 && C::E2 % 2 == 0 && k1 != 0) {
            const float4 *Rq = reinterpret_cast<const float4 *>(
                Rrow + (flip ? (C::E1 - k1) * C::E2 : col_d));
            static_for<0, C::E2 / 2>([&](auto J) {
                constexpr int j = decltype(J)::value;
                const float4 q = Rq[flip ? C::E2 / 2 - 1 - j : j];      // (R is written by this kernel: no ld.nc)
                cf v0 = flip ? cmake(q.z, -q.w) : cmake(q.x, q.y);     // k2 = 2j
                cf v1 = flip ? cmake(q.x, -q.y) : cmake(q.z, q.w);     // k2 = 2j + 1
                x[2 * j] = cmul(x[2 * j], v0);
                x[2 * j + 1] = cmul(x[2 * j + 1], v1);
            });
        } else
#endif
        static_for<0, C::E2>([&](auto K2) {
            constexpr int k2 = decltype(K2)::value;
            cf v = flip ? ((k1 == 0 && k2 == 0) ? Rrow[0] : Rd[-k2]) : Rd[k2];
            if (flip) v.y = -v.y;   // conj(R[ky,kx]) = R[P-ky,P-kx]; storage holds conj(R)
            x[k2] = cmul(x[k2], v);
        });
        RegFFT<C::E2, true>::run(x);
        static_for<0, C::E2>([&](auto N2) {
            constexpr int n2 = decltype(N2)::value;
            srow[k1 * C::E2 + n2] = cmul(x[n2], tw2<C, true>(n2, k1));
        });
    }
}

// X3: E1-point inverse FFT over k1; sample m = E2*n1+n2 is lag (m - shift) - xc,
// kept (at position m of the row buffer) for 0 <= m - shift < nx.
template <class C, bool FAST = false>
SPX_HD void x3_pass(const Ctx<C> &c, int tid, int cb, int ce, int shift) {
    const int warp = warp_of(tid), lane = tid & 31;
    const int nx = FAST ? C::NMAX : c.ps->nx;
    SPX_UNROLL_BIG
    for (int wi = cb * C::E2 + warp; wi < ce * C::E2; wi += C::NW) {
        const int chunk = wi / C::E2, n2 = wi - chunk * C::E2;
        const int r = row_of<C, false>(c, chunk, lane);
        if (C::INPLACE && r < 0) continue;
        const cf *srow = srow_of<C>(c, r, lane);
        cf x[C::E1];
        static_for<0, C::E1>([&](auto K1) {
            constexpr int k1 = decltype(K1)::value;
            x[k1] = srow[k1 * C::E2 + n2];
        });
        RegFFT<C::E1, true>::run(x);
        if (r >= 0) {
            // DIRECT: lag sample l = m - shift goes to column l of the W row
            cf *out = C::DIRECT ? c.W + r * C::WP - shift : srow_of<C>(c, r, lane);
            static_for<0, C::E1>([&](auto N1) {
                constexpr int n1 = decltype(N1)::value;
                const int m = C::E2 * n1 + n2;
                if (unsigned(m - shift) < unsigned(nx)) out[m] = x[n1];
            });
        }
    }
}

// --------------------------------------------------------------------------
// peak stage: numpy.argmax semantics incl. NaN lattices, then
// centroid.find_peak with peak_fit_box=5 (cc.py:86) and cc.py:89-93.
// The surface is at least 6x6, so the fit box is always a full 5x5 box
// (centroid.py:165-184) and the least-squares solution has the closed form
// of SURVEY.md section 3.5.  25 lanes fetch one sample each, the six moment
// sums are reduced across the warp, one lane finishes.
// --------------------------------------------------------------------------
struct PeakGeom { int idx, imax, jmax, x1, y1, mode; };   // mode 0: fit, 1: integer (NaN/edge)
struct FitSums { double s0, su, sv, suv, suu, svv; };
SPX_HD FitSums fit_zero() { FitSums f; f.s0 = f.su = f.sv = f.suv = f.suu = f.svv = 0.0; return f; }
SPX_HD void fit_add(FitSums &a, const FitSums &b) {
    a.s0 += b.s0; a.su += b.su; a.sv += b.sv; a.suv += b.suv; a.suu += b.suu; a.svv += b.svv;
}

SPX_HD PeakGeom peak_geometry(const PairState &ps) {
    PeakGeom g;
    const int h2 = 2 * ps.ny, w2 = 2 * ps.nx;
    const int st = ps.status;
    const bool norm = ps.cc_type != SPX_CC;
    const int nf_all = SPX_S_NONFINITE_REF | SPX_S_BAD_NORM;
    const int nf_ph = SPX_S_NONFINITE_P00 | SPX_S_NONFINITE_P10 |
                      SPX_S_NONFINITE_P01 | SPX_S_NONFINITE_P11;
    g.idx = ps.best.idx;
    bool nan = false;
    if ((st & nf_all) || (norm && (st & nf_ph))) {
        g.idx = 0; nan = true;                      // whole surface NaN
    } else if (st & nf_ph) {                        // CC: first NaN lattice, row-major
        g.idx = (st & SPX_S_NONFINITE_P00) ? 0 : (st & SPX_S_NONFINITE_P10) ? 1
              : (st & SPX_S_NONFINITE_P01) ? w2 : w2 + 1;
        nan = true;
    }
    g.jmax = g.idx / w2; g.imax = g.idx - g.jmax * w2;
    // centroid.py:165-184 for a 5x5 box inside a >= 6x6 image
    g.x1 = g.imax - 2; if (g.x1 < 0) g.x1 = 0; if (g.x1 > w2 - 5) g.x1 = w2 - 5;
    g.y1 = g.jmax - 2; if (g.y1 < 0) g.y1 = 0; if (g.y1 > h2 - 5) g.y1 = h2 - 5;
    // every 5x5 box around a NaN peak contains NaN -> integer position; a peak in
    // row/column 0 sits on the box edge (centroid.py:171-172) -> integer position
    g.mode = (nan || g.imax == 0 || g.jmax == 0) ? 1 : 0;
    if (nan) g.mode |= 2;
    return g;
}

template <class LoadF>
SPX_HD FitSums peak_lane(const PairState &ps, const PeakGeom &g, int lane, LoadF ldicc) {
    FitSums f = fit_zero();
    if (g.mode == 0 && lane < 25) {
        const int dj = lane / 5, di = lane - 5 * dj;
        const double d = double(ldicc((g.y1 + dj) * (2 * ps.nx) + g.x1 + di));
        const double u = double(di - 2), v = double(dj - 2);
        f.s0 = d; f.su = u * d; f.sv = v * d; f.suv = u * v * d; f.suu = u * u * d; f.svv = v * v * d;
    }
    return f;
}

SPX_HD void peak_final(const PairState &ps, const PeakGeom &g, const FitSums &f,
                       double *shift2, int *peak, int *status) {
    const int h2 = 2 * ps.ny, w2 = 2 * ps.nx;
    int st = ps.status & ~kSawNonFinite;
    double x = double(g.imax), y = double(g.jmax);
    if (g.mode != 0) {
        st |= (g.imax == 0 || g.jmax == 0) ? SPX_S_PEAK_EDGE : SPX_S_FIT_NONFINITE;
    } else {
        const double c10 = f.su / 50.0, c01 = f.sv / 50.0, c11 = f.suv / 100.0;
        const double c20 = (f.suu - 2.0 * f.s0) / 70.0, c02 = (f.svv - 2.0 * f.s0) / 70.0;
        const double det = 4.0 * c02 * c20 - c11 * c11;                 // centroid.py:216
        if (det <= 0.0 || ((c20 > 0.0 && c02 >= 0.0) || (c20 >= 0.0 && c02 > 0.0))) {
            x = double(g.x1) + 2.5; y = double(g.y1) + 2.5;             // centroid.py:217-225
            st |= SPX_S_NO_MAXIMUM;
        } else {
            const double xm = (c01 * c11 - 2.0 * c02 * c10) / det + double(g.x1 + 2);
            const double ym = (c10 * c11 - 2.0 * c01 * c20) / det + double(g.y1 + 2);
            if (xm > 0.0 && xm < double(w2) - 1.0 && ym > 0.0 && ym < double(h2) - 1.0) {
                x = xm; y = ym;                                          // centroid.py:230-231
            } else {
                st |= (xm != xm || ym != ym) ? SPX_S_FIT_NONFINITE : SPX_S_FIT_OUTSIDE;
            }
        }
    }
    const int xc4 = (w2 - 1) / 4, yc4 = (h2 - 1) / 4;                   // cc.py:89-90
    shift2[0] = 0.5 * x - double(xc4);                                   // cc.py:92-93
    shift2[1] = 0.5 * y - double(yc4);
    if (peak) {
        *peak = g.idx;
    } else {
        // packed result table (dx, dy, peak, status) as four doubles: the row the
        // multi-GPU all-gather sends (ints up to 2^31 are exact in float64)
        shift2[2] = double(g.idx);
        shift2[3] = double(st & ~kSawNonFinite);
    }
    *status = st & ~kSawNonFinite;
}

// value the reference's surface has at flat index idx when inputs were
// non-finite / the normalisation failed (used to patch a requested ICC)
SPX_HD bool icc_is_nan(const PairState &ps, int status, int idx) {
    const int w2 = 2 * ps.nx;
    const int j = idx / w2, i = idx - j * w2;
    const int lattice = (j & 1) * 2 + (i & 1);         // 0:00 1:10 2:01 3:11
    const bool norm = ps.cc_type != SPX_CC;
    const int nf_ph = SPX_S_NONFINITE_P00 | SPX_S_NONFINITE_P10 |
                      SPX_S_NONFINITE_P01 | SPX_S_NONFINITE_P11;
    if (status & (SPX_S_NONFINITE_REF | SPX_S_BAD_NORM)) return true;
    if (norm && (status & nf_ph)) return true;
    return (status & (SPX_S_NONFINITE_P00 << lattice)) != 0;
}

// What the kernel knows about one pair before any pixel is read.
struct PairJob {
    const float *base[5];
    long long pitch;
    int H, W, x1, y1, ny, nx, cc_type, status0;
    unsigned flags;
    bool patch_icc;        // ICC is a user-visible output: write NaN lattices
};

// Per-pair state, filled by six threads in parallel (part 0..4: tile slots,
// part 5: scalars).
SPX_HD void fill_state(PairState &ps, const PairJob &job, int part) {
    if (part < 5) {
        ps.base[part] = job.base[part];
        ps.pitch[part] = job.pitch;
        ps.y0[part] = job.y1; ps.x0[part] = job.x1;
        return;
    }
    ps.H = job.H; ps.W = job.W;
    ps.ny = job.ny; ps.nx = job.nx;
    ps.yc = (job.ny - 1) / 2; ps.xc = (job.nx - 1) / 2;
    ps.cc_type = job.cc_type; ps.flags = job.flags;
    ps.mu = 0.f; ps.mu_ref = 0.f; ps.rscale = 1.f;
    ps.keep = job.patch_icc ? 0 : 1; ps.in_w = -1; ps.ysh = 0; ps.cout = 0;
    ps.ylo = job.y1 < 0 ? -job.y1 : 0;
    ps.yhi = job.y1 + job.ny > job.H ? job.H - job.y1 : job.ny;
    ps.xlo = job.x1 < 0 ? -job.x1 : 0;
    ps.xhi = job.x1 + job.nx > job.W ? job.W - job.x1 : job.nx;
    ps.interior = job.x1 >= 0 && job.y1 >= 0 && job.x1 + job.nx <= job.W &&
                  job.y1 + job.ny <= job.H;
    ps.report = ((job.flags & SPX_F_ZERO_BAD_REF) ? 0 : SPX_S_NONFINITE_REF) |
                ((job.flags & SPX_F_ZERO_BAD_PHASES) ? 0 :
                 (SPX_S_NONFINITE_P00 | SPX_S_NONFINITE_P10 | SPX_S_NONFINITE_P01 |
                  SPX_S_NONFINITE_P11));
    ps.status = job.status0;
    ps.best = best_init();
}

// --------------------------------------------------------------------------
// fit samples of a surface that was kept in W (SurfaceInW)
// --------------------------------------------------------------------------
// part t of kWinN*kWinN: copy sample (kc - kWinR + t / kWinN, lc - kWinR + t % kWinN) of the
// phase pair now in W into window w
template <class C>
SPX_HD void save_window(const Ctx<C> &c, int w, int kc, int lc, int t, int ysh, int cout) {
    const PairState &ps = *c.ps;
    const int dk = t / kWinN, dl = t - dk * kWinN;
    const int k = kc - kWinR + dk, l = lc - kWinR + dl;
    if (t == 0) { c.win->k0[w] = kc - kWinR; c.win->l0[w] = lc - kWinR; }
    if (k >= 0 && k < ps.ny && l >= 0 && l < ps.nx)
        c.win->v[w][dk][dl] = c.W[(k + ysh) * C::WP + l + cout];
}
// surface sample at flat index idx of the (2ny, 2nx) surface
template <class C>
SPX_HD float kept_sample(const Ctx<C> &c, int idx) {
    const PairState &ps = *c.ps;
    const int w2 = 2 * ps.nx;
    const int jj = idx / w2, ii = idx - jj * w2;
    const int b = jj & 1, k = jj >> 1, l = ii >> 1;
    const cf v = (b == ps.in_w) ? c.W[(k + ps.ysh) * C::WP + l + ps.cout]
                                : c.win->v[b][k - c.win->k0[b]][l - c.win->l0[b]];
    return (ii & 1) ? v.y : v.x;
}
// do the samples of phase pair b that the fit box needs lie inside window b?
template <class C>
SPX_HD bool window_covers(const Ctx<C> &c, const PeakGeom &g, int b) {
    if (g.mode != 0) return true;                       // no fit, no samples
    const int jlo = g.y1 + ((g.y1 & 1) != b), jhi = g.y1 + 4 - (((g.y1 + 4) & 1) != b);
    const int klo = jlo >> 1, khi = jhi >> 1, llo = g.x1 >> 1, lhi = (g.x1 + 4) >> 1;
    const int k0 = c.win->k0[b], l0 = c.win->l0[b];
    return klo >= k0 && khi < k0 + kWinN && llo >= l0 && lhi < l0 + kWinN;
}

// --------------------------------------------------------------------------
// reference transform + the two phase pairs (see the data flow at the top)
// --------------------------------------------------------------------------
#ifdef SPX_LAZY_NODETECT                 // developer knob: cost of the detector (results unchecked)
constexpr int kLazyDetect = 0;
#else
constexpr int kLazyDetect = 2;
#endif

// One phase pair (round IT: 0 / 1 = first / second pair, 2 = the first pair once more):
// C1 C2 X1 X2 X3 IC1 IC2 + argmax.  Returns false when the statistics-free attempt has to
// be abandoned (lazy_decide).
template <class C, bool FAST, bool CHECK, class Ex>
SPX_HD bool phase_pair(Ctx<C> &c, Ex &ex, const int IT, bool lazy, bool keep, int yoff, int xoff,
                       int ysh, int xsh, int cin, int cout) {
    PairState *ps = c.ps;
    const int b = IT & 1;
    const int lz = (lazy && IT < 2) ? (b == 0 ? kLazyDetect : 1) : 0;
    [[maybe_unused]] const int nch = ((FAST ? C::NMAX : ps->nx) + 31) >> 5;    // 32-column chunks
    if constexpr (C::FUSE) {
        // C1 -> C2 -> slab, chunk by chunk through S (W is written once, finished)
        SPX_NO_UNROLL
        for (int j = 0; j < nch; ++j) {
            ex.par([&](int tid) {
                const int fl = c1_pass<C, FAST, false, CHECK, true>(c, tid, 1 + 2 * b, 2 + 2 * b, yoff,
                                                                    cin, lz, j);
                if (fl) ex.or_status(&ps->status, fl);
            });
#ifndef SPX_NO_DIRECT_W
            ex.par([&](int tid) { c2_pass<C, FAST, true, true, false>(c, tid, cin, j); });
#else
            ex.par([&](int tid) { c2_pass<C, FAST, true>(c, tid, cin, j); });
            ex.par([&](int tid) { chunk_out<C, false>(c, tid, j, cin); });
#endif
        }
    } else {
        ex.par([&](int tid) {
            const int fl = c1_pass<C, FAST, false, CHECK>(c, tid, 1 + 2 * b, 2 + 2 * b, yoff, cin, lz);
            if (fl) ex.or_status(&ps->status, fl);
        });
    }
    // my 32-column chunks (the tile columns this CTA loaded): [m0, m1)
    [[maybe_unused]] const int m0 = c.rank * (C::NXL / 32), m1 = m0 + C::NXL / 32;
    if (C::SPLIT) {
        ex.cluster_arrive();                // my stores into the peers' rows are on their way
        ex.par([&](int tid) { c2_pass<C, FAST>(c, tid, cin, 0, m0, m1); });
        ex.cluster_wait();                  // the peers' part of my rows has arrived
    } else if (C::CL > 1) ex.cluster_sync();
    // (the detector keys of both CTAs are complete after this barrier; every thread
    // evaluates the same few comparisons, so the branch is uniform)
    if (lazy && IT < 2 && !lazy_decide(c.xl->det[c.dslot], b == 0)) return false;
    // staged tiles: every thread is past its last tile read (the barrier that ended this C1
    // pass) and the pair will not be restarted any more: the next pair's copy may overwrite them
    if (C::STAGE && IT == 1) ex.tiles_release();
    if (C::SPLIT) ex.par([&](int tid) {
        c2_pass<C, FAST>(c, tid, cin, 0, 0, m0);
        c2_pass<C, FAST>(c, tid, cin, 0, m1);
    });
    else if (!C::FUSE) ex.par([&](int tid) { c2_pass<C, FAST>(c, tid, cin); });
    if (C::INPLACE) {
        ex.par([&](int tid) { x1_pass<C, FAST, false>(c, tid, 0, C::NCHUNK, xoff); });
        ex.par([&](int tid) { x2_pass<C>(c, tid, 0, C::NCHUNK); });
        ex.par([&](int tid) { x3_pass<C, FAST>(c, tid, 0, C::NCHUNK, xsh); });
    } else {
        for (int k = 0; k < C::NCHUNK; ++k) {
            if (!C::DIRECT) ex.par([&](int tid) { stage_in<C, false>(c, tid, k, xoff); });
            ex.par([&](int tid) {
                if (C::FUSE) prefetch_row_batch<C, false>(c, tid, k + 1);
                x1_pass<C, FAST, false>(c, tid, k, k + 1, xoff);
            });
            ex.par([&](int tid) { x2_pass<C>(c, tid, k, k + 1); });
            ex.par([&](int tid) { x3_pass<C, FAST>(c, tid, k, k + 1, xsh); });
            if (!C::DIRECT) ex.par([&](int tid) { stage_out<C>(c, tid, k, xsh); });
        }
    }
    // W now holds this pair's surface (keep).  After the first pair the warp that finishes
    // the argmax also saves the window around the best sample of that pair, before the
    // barrier that lets the next C1 pass overwrite W.
    auto after = [&](int lane) {
        if (!keep) return;
        if (lane == 0) { ps->in_w = b; ps->ysh = ysh; ps->cout = cout; }
        if (IT != 0) return;
        const int kc = (ps->best.idx / (2 * ps->nx)) >> 1, lc = (ps->best.idx % (2 * ps->nx)) >> 1;
        for (int t = lane; t < kWinN * kWinN; t += 32) save_window<C>(c, 0, kc, lc, t, ysh, cout);
    };
    if constexpr (C::FUSE) {
        // slab -> IC1 -> IC2, chunk by chunk through S; every thread keeps its running best
        // over the chunks (the merge is order independent: larger value, then smaller index)
        Best acc = best_init();
        SPX_NO_UNROLL
        for (int j = 0; j < nch; ++j) {
#ifndef SPX_NO_DIRECT_W
            ex.par([&](int tid) {
                prefetch_col_chunk<C>(c, tid, j + 1, cout);
                ic1_pass<C, FAST, true, true>(c, tid, cout, j);
            });
#else
            ex.par([&](int tid) { chunk_in<C>(c, tid, j, cout); });
            ex.par([&](int tid) {
                prefetch_col_chunk<C>(c, tid, j + 1, cout);
                ic1_pass<C, FAST, true>(c, tid, cout, j);
            });
#endif
            ex.par([&](int tid) {
                const Best t = ic2_pass<C, FAST, true>(c, tid, b, ysh, cout, j);
                if (best_better(t, acc)) acc = t;
            });
        }
        ex.reduce_best_into(&ps->best, [&](int) { return acc; }, after);
    } else {
        if (C::SPLIT) {
            ex.par([&](int tid) {           // the columns the peers' IC2 will read from my rows
                ic1_pass<C, FAST>(c, tid, cout, 0, 0, m0);
                ic1_pass<C, FAST>(c, tid, cout, 0, m1);
            });
            ex.cluster_arrive();
            ex.par([&](int tid) { ic1_pass<C, FAST>(c, tid, cout, 0, m0, m1); });
            ex.cluster_wait();              // the peers' rows hold the lags of my columns
        } else {
            ex.par([&](int tid) { ic1_pass<C, FAST>(c, tid, cout); });
            if (C::CL > 1) ex.cluster_sync();   // the peer's rows hold their lags: IC2 may read them
        }
        ex.reduce_best_into(&ps->best, [&](int tid) { return ic2_pass<C, FAST>(c, tid, b, ysh, cout); },
                            after);
    }
    // the peer must not start writing my rows (its next C1) while I still read
    // them; after the second pair the exchange barrier of the peak stage does that
    if (C::CL > 1 && b == 0) ex.cluster_sync();
    return true;
}

// lazy: statistics-free NCC attempt; returns false (after a C1 pass over image tiles) when
// the detector could not prove the normalisation well defined.

template <class C, bool FAST, bool CHECK, class Ex>
SPX_HD bool run_transforms(Ctx<C> &c, Ex &ex, bool lazy) {
    PairState *ps = c.ps;
    // placement offset of the image tiles and the resulting index shift of the
    // wanted lags: sample m of the circular correlation is lag (m - off)
    const int yoff = FAST ? C::NMAX / 2 : ps->yc, xoff = FAST ? C::NMAX / 2 : ps->xc;
    // (compile-time in both cases: the generic path places the tiles at (yc, xc) itself)
    constexpr int kSh = FAST ? C::NMAX / 2 - (C::NMAX - 1) / 2 : 0;
    const int ysh = kSh, xsh = kSh;
    // INPLACE: W columns are padded positions (input column i sits at i + xoff,
    // wanted lag sample l at l + xsh)
    const int cin = C::INPLACE ? xoff : 0, cout = C::INPLACE ? xsh : 0;

    // ---- reference image -> conj(R) * scale, half spectrum ---------------
#ifdef SPX_LAZY_NOREF                        // developer knob: cost of the reference-side detector
    const int lzr = lazy ? 1 : 0;
#else
    const int lzr = lazy ? kLazyDetect : 0;
#endif
    if constexpr (C::FUSE) {
        const int nch = ((FAST ? C::NMAX : ps->nx) + 31) >> 5;
        SPX_NO_UNROLL
        for (int j = 0; j < nch; ++j) {
            ex.par([&](int tid) {
                const int fl = c1_pass<C, FAST, true, CHECK, true>(c, tid, 0, -1, 0, 0, lzr, j);
                if (fl) ex.or_status(&ps->status, fl);
            });
#ifndef SPX_NO_DIRECT_W
            ex.par([&](int tid) { c2_pass<C, FAST, true, true, true>(c, tid, 0, j); });
#else
            ex.par([&](int tid) { c2_pass<C, FAST, true>(c, tid, 0, j); });
            ex.par([&](int tid) { chunk_out<C, true>(c, tid, j, 0); });
#endif
        }
    } else {
        ex.par([&](int tid) {
            const int fl = c1_pass<C, FAST, true, CHECK>(c, tid, 0, -1, 0, 0, lzr);
            if (fl) ex.or_status(&ps->status, fl);
        });
        if (C::SPLIT) {
            const int m0 = c.rank * (C::NXL / 32), m1 = m0 + C::NXL / 32;
            ex.cluster_arrive();
            ex.par([&](int tid) { c2_pass<C, FAST>(c, tid, 0, 0, m0, m1); });
            ex.cluster_wait();
            ex.par([&](int tid) {
                c2_pass<C, FAST>(c, tid, 0, 0, 0, m0);
                c2_pass<C, FAST>(c, tid, 0, 0, m1);
            });
        } else {
            if (C::CL > 1) ex.cluster_sync();   // the peer's half of my rows has arrived
            ex.par([&](int tid) { c2_pass<C, FAST>(c, tid, 0); });
        }
    }
    if (C::INPLACE) {
        ex.par([&](int tid) { x1_pass<C, FAST, true>(c, tid, 0, C::RCHUNK, 0); });
        ex.par([&](int tid) { x2_ref_pass<C>(c, tid, 0, C::RCHUNK); });
    } else {
        for (int b = 0; b < C::RCHUNK; ++b) {
            if (!C::DIRECT) ex.par([&](int tid) { stage_in<C, true>(c, tid, b, 0); });
            ex.par([&](int tid) {
                if (C::FUSE) prefetch_row_batch<C, true>(c, tid, b + 1);
                x1_pass<C, FAST, true>(c, tid, b, b + 1, 0);
            });
            ex.par([&](int tid) { x2_ref_pass<C>(c, tid, b, b + 1); });
        }
    }
    if (C::CL > 1) ex.cluster_sync();           // I am done with W: the peer's next C1 may write it

    // ---- phase pairs: (im00, im10) -> even rows, (im01, im11) -> odd rows --
    // Round 0 and 1 are the two pairs - ONE loop body with a runtime round (two specialised
    // copies spill in the 64 px class and cost 7 %); round 2 repeats the first pair when
    // its fit samples were not kept (see SurfaceInW) - rare, a cold second copy.
    const bool keep = SurfaceInW<C>::value && ps->keep != 0;
#if !defined(SPX_HOST_EMU)
#pragma unroll 1
#endif
    for (int it = 0; it < 2; ++it)
        if (!phase_pair<C, FAST, CHECK>(c, ex, it, lazy, keep, yoff, xoff, ysh, xsh, cin, cout))
            return false;
    if constexpr (SurfaceInW<C>::value) {
        if (keep) {
            const PeakGeom g = peak_geometry(*ps);      // ps->best is final: uniform decision
            if (!window_covers<C>(c, g, 0)) {
#ifdef SPX_HOST_EMU
                if (getenv("SPX_EMU_COUNT")) fprintf(stderr, "fallback: third round\n");
#endif
                // keep the second pair's samples around the global best, then redo the first
                ex.parts(kWinN * kWinN, [&](int t) {
                    save_window<C>(c, 1, g.jmax >> 1, g.imax >> 1, t, ysh, cout);
                });
                phase_pair<C, FAST, CHECK>(c, ex, 2, false, keep, yoff, xoff, ysh, xsh, cin, cout);
            }
        }
    }
    return true;
}

// --------------------------------------------------------------------------
// The pass sequence for ONE pair.  `Ex` supplies the execution model:
//   par(f)          run f(tid) for every thread of the CTA, then barrier
//   single(f)       run f() on one thread, then barrier
//   parts(n, f)     run f(0..n-1) on n different threads, then barrier
//   reduce_stats(f) block reduction of f(tid); meaningful in the next single()
//   reduce_best_into(dst, f, after)  block argmax of f(tid) merged into *dst, then
//                   after(lane) on the 32 lanes of the warp that merged it, barrier
//   warp_fit(lane_f, final_f) one warp: sum lane_f(lane) over 32 lanes, then
//                   final_f(sum) on one lane, then barrier
//   cluster_sync()  barrier over all CTAs of a cluster (CL > 1 only)
//   cluster_arrive() / cluster_wait()  the two halves of such a barrier (split phase)
//   tiles_wait()    (STAGE) block until the staged tiles of the current pair have landed
//   tiles_release() (STAGE) the tiles were read for the last time: start staging the next pair
//   global_fence()  make this thread block's global writes visible to the peer
//   or_status(p,v)  atomic OR
// The device executor is in spx_cc.cu; tools/host_emu.cu runs the same
// sequence serially on the CPU for debugging.
// --------------------------------------------------------------------------
template <class C, class Ex>
SPX_HD void process_pair(Ctx<C> &c, Ex &ex, const PairJob &job,
                         double *shift2, int *peak, int *status) {
    PairState *ps = c.ps;
    const bool fast = job.ny == C::NMAX && job.nx == C::NMAX && job.x1 >= 0 && job.y1 >= 0 &&
                      job.x1 + job.nx <= job.W && job.y1 + job.ny <= job.H;
    const int nonfinite = SPX_S_NONFINITE_REF | SPX_S_NONFINITE_P00 | SPX_S_NONFINITE_P10 |
                          SPX_S_NONFINITE_P01 | SPX_S_NONFINITE_P11;
    // NCC without a requested surface: first try without the statistics sweep (lazy_decide).
    // ONE loop body serves both attempts: a second inlined copy of the transforms would
    // not fit the instruction cache.
    bool lazy = job.cc_type == SPX_NCC && !job.patch_icc && !(job.flags & SPX_F_EXACT_NORM);
#ifdef SPX_NO_LAZY                       // developer knob (tools/variant.sh): A/B of the sweep
    lazy = false;
#endif
#if !defined(SPX_HOST_EMU)
#pragma unroll 1
#endif
    for (;;) {
        if (C::STAGE) ex.tiles_wait();          // the TMA has landed this pair's tiles
        ex.parts(7, [&](int part) {
            if (part < 6) { fill_state(*ps, job, part); return; }
            // single CTA: nobody else touches the detector slot (cluster mode: see Xchg)
            if (C::CL == 1) { c.xl->det[c.dslot][0] = 0u; c.xl->det[c.dslot][1] = 0u; }
        });

        // ---- normalisation statistics (cc.py:131-156) ----------------------
        if (job.cc_type != SPX_CC && !lazy) {
            StatPartial tot = ex.reduce_stats([&](int tid) { return stats_thread<C>(c, tid); });
            if (C::CL > 1) {
                // each CTA scanned its share of the columns: exchange and add in a fixed
                // order so that all CTAs derive bit-identical statistics
                ex.single([&]() {
                    const double v[8] = {tot.cnt, tot.sum, tot.sq, tot.rcnt, tot.rsum, tot.rsq,
                                         double(tot.flags), 0.0};
                    for (int rk = 0; rk < C::CL; ++rk) {
                        Xchg *x = X_of<C>(c, rk);
                        for (int k = 0; k < 8; ++k) x->stat[c.rank][k] = v[k];
                    }
                });
                ex.cluster_sync();
                ex.single([&]() {
                    StatPartial t = stat_zero();
                    for (int rk = 0; rk < C::CL; ++rk) {
                        const double *v = c.xl->stat[rk];
                        t.cnt += v[0]; t.sum += v[1]; t.sq += v[2];
                        t.rcnt += v[3]; t.rsum += v[4]; t.rsq += v[5]; t.flags |= int(v[6]);
                    }
                    stats_finalize(*ps, t);
                });
            } else {
                ex.single([&]() { stats_finalize(*ps, tot); });
            }
        }

        // tiles already scanned by the statistics pass and found finite -> no NaN tracking
        const bool check = lazy || job.cc_type == SPX_CC ||
                           (ps->status & (nonfinite | kSawNonFinite)) != 0;
        bool done;
        if constexpr (C::HAS_FAST) {
            if (fast) {
                if (check) done = run_transforms<C, true, true>(c, ex, lazy);
                else done = run_transforms<C, true, false>(c, ex, false);
            } else {
                done = run_transforms<C, false, true>(c, ex, lazy);
            }
        } else {
            done = run_transforms<C, false, true>(c, ex, lazy);
        }
        if (done) break;
        lazy = false;                   // degenerate or extreme input: the exact path decides
        // in cluster mode the peer may still be inside its detector pass; nothing it
        // writes from now on (its own W rows) is touched before the next cluster barrier
    }

    // ---- argmax over the interlaced surface + 5x5 quadratic fit ----------
    if (C::CL > 1) {
        // hand my best sample and my status bits to every CTA of the cluster
        ex.parts(C::CL, [&](int rk) {
            Xchg *x = X_of<C>(c, rk);
            x->bestv[c.rank] = ps->best.v;
            x->besti[c.rank] = ps->best.idx;
            x->status[c.rank] = ps->status;
            if (rk == 0) { c.xl->det[c.dslot ^ 1][0] = 0u; c.xl->det[c.dslot ^ 1][1] = 0u; }   // next pair's slot
        });
        // the cluster barrier is a release/acquire fence at cluster scope: my part of
        // the surface (global memory) is visible to rank 0 after it
        ex.cluster_sync();
        ex.single([&]() {
            for (int rk = 0; rk < C::CL; ++rk) {
                if (rk == c.rank) continue;
                Best o; o.v = c.xl->bestv[rk]; o.idx = c.xl->besti[rk];
                if (best_better(o, ps->best)) ps->best = o;
                ps->status |= c.xl->status[rk];
            }
        });
    }
    if (C::CL == 1 || c.rank == 0) {
        const float *icc = c.icc;
        ex.warp_fit(
            [&](int lane) {
                const PeakGeom g = peak_geometry(*ps);
                return peak_lane(*ps, g, lane, [&](int i) {
                    if (SurfaceInW<C>::value && ps->keep) return kept_sample<C>(c, i);
                    return ex.load_icc(icc + i);
                });
            },
            [&](const FitSums &f) {
                const PeakGeom g = peak_geometry(*ps);
                int st;
                peak_final(*ps, g, f, shift2, peak, &st);
                if (status) *status = st;
                ps->status = st;
            });
    }
    if (job.patch_icc && (C::CL == 1 || c.rank == 0)) {
        ex.par([&](int tid) {
            const int st = ps->status;
            const int nanbits = SPX_S_NONFINITE_REF | SPX_S_BAD_NORM | SPX_S_NONFINITE_P00 |
                                SPX_S_NONFINITE_P10 | SPX_S_NONFINITE_P01 | SPX_S_NONFINITE_P11;
            if (st & nanbits) {
                const int ne = 4 * ps->ny * ps->nx;
                union { unsigned u; float f; } qnan; qnan.u = 0x7fc00000u;
                for (int i = tid; i < ne; i += C::T)
                    if (icc_is_nan(*ps, st, i)) c.icc[i] = qnan.f;
            }
        });
    }
}

// my rows of W (local index) whose frequency lies in the stored half ky <= P/2;
// returns their number
template <class C>
SPX_HD int build_rrow(uint16_t *rrow, int rank) {
    int k = 0;
    for (int lr = 0; lr < C::PL; ++lr)
        if (ky_of_row<C>(lr, rank) <= C::P / 2) rrow[k++] = uint16_t(lr);
    return k;
}

}  // namespace spx
