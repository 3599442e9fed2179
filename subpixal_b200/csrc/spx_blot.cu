// (6) Batched blot of drizzled cutouts onto image-cutout grids at several
// sub-pixel phases (reference blot.py:79-155 `blot_cutout` -> drizzlepac
// `cdriz.tblot`, interp='poly5', misval=0; called four times per source at
// align.py:661-679 with the target grid displaced by half a pixel).
//
// Per output pixel: target pixel (i + ddx, j + ddy) -> source position through a
// per-source biquadratic map (float64, rounded once to float32 as drizzle's
// pixmap is), range test against the source cutout, 6x6 quintic interpolation in
// Everett's central-difference form with the linear edge extension
// 2*f(edge) - f(mirror), flux scale.  All float32 arithmetic uses explicit
// round-to-nearest intrinsics (no FMA contraction) in the operation order of
// oracle/blot_oracle.py, so results are bit-identical to that restatement.
//
// Work decomposition: blockIdx.y = source, all phases of one source are produced by
// the same CTAs (its source cutout stays in L1), lanes run along x so stores are
// coalesced and the 36 loads of neighbouring lanes hit the same lines.
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/subpixal_b200.h"

namespace spx {

constexpr int kBlotThreads = 256;
constexpr int kMaxPhases = 8;
#ifndef SPX_BLOT_SMEM_TABLE_MAX
#define SPX_BLOT_SMEM_TABLE_MAX (44 * 1024)      // >= 5 CTAs per SM (sources up to about 50 px); larger tables are faster through L2
#endif

struct BlotDev {
    const float *src; const long long *src_off; const int *src_shape;
    const double *map; const float *scale;
    const int *tgt_shape; const long long *out_off;
    float *out;
    float4 *ws; const long long *ws_off;      // prepared sources (poly5), see blot_prep_kernel
    int n, nph, interp;
    float misval;
    double ph[kMaxPhases][2];
};

__device__ __forceinline__ float fa(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fs(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fm(float a, float b) { return __fmul_rn(a, b); }

// quintic through six equally spaced samples (positions -2..3) at 0 <= s < 1
__device__ __forceinline__ float everett(float fm2, float fm1, float f0, float f1, float f2,
                                         float f3, float s) {
    const float s2 = fm(s, s), t = fs(1.f, s), t2 = fm(t, t);
    const float cd20 = fm(1.0f / 6.0f, fa(fs(f1, fm(2.f, f0)), fm1));
    const float cd21 = fm(1.0f / 6.0f, fa(fs(f2, fm(2.f, f1)), f0));
    const float cd40 = fm(1.0f / 120.0f,
                          fa(fs(fa(fs(fm2, fm(4.f, fm1)), fm(6.f, f0)), fm(4.f, f1)), f2));
    const float cd41 = fm(1.0f / 120.0f,
                          fa(fs(fa(fs(fm1, fm(4.f, f0)), fm(6.f, f1)), fm(4.f, f2)), f3));
    const float a = fm(s, fa(f1, fm(fs(s2, 1.f), fa(cd21, fm(fs(s2, 4.f), cd41)))));
    const float b = fm(t, fa(f0, fm(fs(t2, 1.f), fa(cd20, fm(fs(t2, 4.f), cd40)))));
    return fa(a, b);
}

struct Fold { int m, e; bool out; };
__device__ __forceinline__ Fold fold(int k, int n) {
    Fold f;
    const bool lo = k < 0, hi = k >= n;
    int m = lo ? -k : (hi ? 2 * (n - 1) - k : k);
    m = m < 0 ? 0 : (m > n - 1 ? n - 1 : m);
    f.m = m; f.e = lo ? 0 : (hi ? n - 1 : k); f.out = lo || hi;
    return f;
}

// source sample with the linear edge extension, columns first, then rows
__device__ __forceinline__ float ext_sample(const float *d, int h, int w, int j, int i) {
    const Fold fx = fold(i, w), fy = fold(j, h);
    auto col = [&](int row) {
        const float direct = __ldg(d + (long long)row * w + fx.m);
        return fx.out ? fs(fm(2.f, __ldg(d + (long long)row * w + fx.e)), direct) : direct;
    };
    const float direct = col(fy.m);
    return fy.out ? fs(fm(2.f, col(fy.e)), direct) : direct;
}

__device__ __forceinline__ float poly5(const float *d, int h, int w, float x, float y) {
    const int ix = int(x), iy = int(y);
    const float sx = fs(x, float(ix)), sy = fs(y, float(iy));
    float z[6];
    if (ix >= 2 && ix + 3 < w && iy >= 2 && iy + 3 < h) {           // interior: plain loads
        const float *p = d + (long long)(iy - 2) * w + (ix - 2);
#pragma unroll
        for (int r = 0; r < 6; ++r) {
            float f[6];
#pragma unroll
            for (int c = 0; c < 6; ++c) f[c] = __ldg(p + (long long)r * w + c);
            z[r] = everett(f[0], f[1], f[2], f[3], f[4], f[5], sx);
        }
    } else {
#pragma unroll 1
        for (int r = 0; r < 6; ++r) {
            float f[6];
#pragma unroll
            for (int c = 0; c < 6; ++c) f[c] = ext_sample(d, h, w, iy - 2 + r, ix - 2 + c);
            z[r] = everett(f[0], f[1], f[2], f[3], f[4], f[5], sx);
        }
    }
    return everett(z[0], z[1], z[2], z[3], z[4], z[5], sy);
}

// cubic through four equally spaced samples (positions -1..2) at 0 <= s < 1
__device__ __forceinline__ float everett3(float fm1, float f0, float f1, float f2, float s) {
    const float s2 = fm(s, s), t = fs(1.f, s), t2 = fm(t, t);
    const float cd20 = fm(1.0f / 6.0f, fa(fs(f1, fm(2.f, f0)), fm1));
    const float cd21 = fm(1.0f / 6.0f, fa(fs(f2, fm(2.f, f1)), f0));
    return fa(fm(s, fa(f1, fm(fs(s2, 1.f), cd21))), fm(t, fa(f0, fm(fs(t2, 1.f), cd20))));
}

__device__ __forceinline__ float poly3(const float *d, int h, int w, float x, float y) {
    const int ix = int(x), iy = int(y);
    const float sx = fs(x, float(ix)), sy = fs(y, float(iy));
    float z[4];
    if (ix >= 1 && ix + 2 < w && iy >= 1 && iy + 2 < h) {           // interior: plain loads
        const float *p = d + (long long)(iy - 1) * w + (ix - 1);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            float f[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) f[c] = __ldg(p + (long long)r * w + c);
            z[r] = everett3(f[0], f[1], f[2], f[3], sx);
        }
    } else {
#pragma unroll 1
        for (int r = 0; r < 4; ++r) {
            float f[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) f[c] = ext_sample(d, h, w, iy - 1 + r, ix - 1 + c);
            z[r] = everett3(f[0], f[1], f[2], f[3], sx);
        }
    }
    return everett3(z[0], z[1], z[2], z[3], sy);
}

// ---- prepared sources ------------------------------------------------------------
// The x-direction central differences of Everett's formula belong to the SOURCE
// sample, not to the output pixel: cd2(c) = (E(c+1) - 2E(c) + E(c-1))/6 and
// cd4(c) = (E(c-2) - 4E(c-1) + 6E(c) - 4E(c+1) + E(c+2))/120 with E the edge-
// extended source.  blot_prep_kernel evaluates them once per source pixel on the
// padded domain rows [-2, h+2] x columns [0, w] and stores {E, cd2, cd4, 0} as one
// float4, so the main kernel needs two 128-bit loads and 15 operations per stencil
// row instead of six loads and 41 operations.  Same operations in the same order:
// results are bit-identical to the direct form.
constexpr int kPadRows = 5;      // rows -2 .. h+2
__host__ __device__ __forceinline__ long long prep_elems(int h, int w) {
    return (long long)(h + kPadRows) * (w + 1);
}

// entry t of the prepared table of one source (padded row t / (w+1), column t % (w+1))
__device__ __forceinline__ float4 prep_entry(const float *d, int h, int w, int r, int c) {
    const int j = r - 2;
    float e[5];
    if (j >= 0 && j < h && c >= 2 && c + 2 < w) {                 // interior: plain loads
        const float *p = d + (long long)j * w + (c - 2);
#pragma unroll
        for (int k = 0; k < 5; ++k) e[k] = __ldg(p + k);
    } else {
#pragma unroll
        for (int k = 0; k < 5; ++k) e[k] = ext_sample(d, h, w, j, c - 2 + k);
    }
    float4 o;
    o.x = e[2];
    o.y = fm(1.0f / 6.0f, fa(fs(e[3], fm(2.f, e[2])), e[1]));
    o.z = fm(1.0f / 120.0f, fa(fs(fa(fs(e[0], fm(4.f, e[1])), fm(6.f, e[2])), fm(4.f, e[3])), e[4]));
    o.w = 0.f;
    return o;
}

__global__ void __launch_bounds__(kBlotThreads) blot_prep_kernel(const BlotDev a) {
    const int s = blockIdx.y;
    const int h = a.src_shape[2 * s], w = a.src_shape[2 * s + 1];
    if (h < 1 || w < 1) return;
    const float *d = a.src + a.src_off[s];
    float4 *ws = a.ws + a.ws_off[s];
    const int pw = w + 1, total = (h + kPadRows) * pw;
    for (int t = blockIdx.x * kBlotThreads + threadIdx.x; t < total; t += gridDim.x * kBlotThreads) {
        const int r = t / pw, c = t - r * pw;
        ws[t] = prep_entry(d, h, w, r, c);
    }
}

// one stencil row from the prepared source: samples at columns ix (p0) and ix+1 (p1)
__device__ __forceinline__ float everett_row(float4 p0, float4 p1, float s, float s2m1, float s2m4,
                                             float t, float t2m1, float t2m4) {
    const float a = fm(s, fa(p1.x, fm(s2m1, fa(p1.y, fm(s2m4, p1.z)))));
    const float b = fm(t, fa(p0.x, fm(t2m1, fa(p0.y, fm(t2m4, p0.z)))));
    return fa(a, b);
}

template <bool SMEM>
__device__ __forceinline__ float poly5_prepared(const float4 *ws, int w, float x, float y) {
    const int ix = int(x), iy = int(y);
    const float sx = fs(x, float(ix)), sy = fs(y, float(iy));
    const float s2 = fm(sx, sx), t = fs(1.f, sx), t2 = fm(t, t);
    const float s2m1 = fs(s2, 1.f), s2m4 = fs(s2, 4.f), t2m1 = fs(t2, 1.f), t2m4 = fs(t2, 4.f);
    const float4 *p = ws + (long long)iy * (w + 1) + ix;       // padded row iy = source row iy-2
    float4 q[12];
#pragma unroll
    for (int r = 0; r < 6; ++r) {
        if (SMEM) {                      // table in shared memory: plain (LDS.128) loads
            q[2 * r] = p[r * (w + 1)];
            q[2 * r + 1] = p[r * (w + 1) + 1];
        } else {
            q[2 * r] = __ldg(p + (long long)r * (w + 1));
            q[2 * r + 1] = __ldg(p + (long long)r * (w + 1) + 1);
        }
    }
    float z[6];
#pragma unroll
    for (int r = 0; r < 6; ++r) z[r] = everett_row(q[2 * r], q[2 * r + 1], sx, s2m1, s2m4, t, t2m1, t2m4);
    return everett(z[0], z[1], z[2], z[3], z[4], z[5], sy);
}

__device__ __forceinline__ float bilinear(const float *d, int h, int w, float x, float y) {
    const int ix = int(x), iy = int(y);
    const float sx = fs(x, float(ix)), sy = fs(y, float(iy));
    const float tx = fs(1.f, sx), ty = fs(1.f, sy);
    const int i1 = ix + 1 < w ? ix + 1 : w - 1, j1 = iy + 1 < h ? iy + 1 : h - 1;
    const float d00 = __ldg(d + (long long)iy * w + ix), d10 = __ldg(d + (long long)iy * w + i1);
    const float d01 = __ldg(d + (long long)j1 * w + ix), d11 = __ldg(d + (long long)j1 * w + i1);
    return fa(fa(fa(fm(fm(tx, ty), d00), fm(fm(sx, ty), d10)), fm(fm(tx, sy), d01)),
              fm(fm(sx, sy), d11));
}

__device__ __forceinline__ float nearest(const float *d, int h, int w, float x, float y) {
    int ix = int(fa(x, 0.5f)), iy = int(fa(y, 0.5f));
    ix = ix < w ? ix : w - 1; iy = iy < h ? iy : h - 1;
    return __ldg(d + (long long)iy * w + ix);
}

__device__ __forceinline__ double dma(double a, double b, double c) {   // a*b + c, two roundings
    return __dadd_rn(__dmul_rn(a, b), c);
}

// PREP: 0 = direct evaluation, 1 = prepared table in global memory (blot_prep_kernel),
// 2 = one CTA per source builds the table in its shared memory first (sources small enough:
// no table traffic to and from L2 / HBM, no second launch)
template <int INTERP, int PREP>
__global__ void __launch_bounds__(kBlotThreads) blot_kernel(const BlotDev a) {
    extern __shared__ float4 s_tab[];
    const int s = blockIdx.y;
    const int h = a.src_shape[2 * s], w = a.src_shape[2 * s + 1];
    const int ny = a.tgt_shape[2 * s], nx = a.tgt_shape[2 * s + 1];
    if (h < 1 || w < 1 || ny < 1 || nx < 1) return;
    const float *d = a.src + a.src_off[s];
    float *out = a.out + a.out_off[s];
    const float4 *ws = PREP == 1 ? a.ws + a.ws_off[s] : (PREP == 2 ? s_tab : nullptr);
    // the map's 20 coefficients: shared memory, re-read at the start of every row
    __shared__ double cf[20];
    if (threadIdx.x < 20) cf[threadIdx.x] = __ldg(a.map + 20 * (long long)s + threadIdx.x);
    if (PREP == 2) {
        const int pw = w + 1, total = (h + kPadRows) * pw;
        for (int t = threadIdx.x; t < total; t += kBlotThreads) {
            const int r = t / pw, c = t - r * pw;
            s_tab[t] = prep_entry(d, h, w, r, c);
        }
    }
    __syncthreads();
    const double xc = cf[0], yc = cf[1];
    const float scale = a.scale ? a.scale[s] : 1.f;
    const float xmax = float(w - 1), ymax = float(h - 1);
    // items = (phase, row, 32-pixel segment); a warp owns one item at a time
    const int segs = (nx + 31) >> 5;
    const int rows = a.nph * ny;
    const int lane = threadIdx.x & 31;
    const int wid = blockIdx.x * (kBlotThreads / 32) + (threadIdx.x >> 5);
    const int wstride = gridDim.x * (kBlotThreads / 32);
    for (int rowph = wid; rowph < rows; rowph += wstride) {
        const int k = rowph / ny, j = rowph - k * ny;
        const double phx = a.ph[k][0];
        const double v = __dsub_rn(__dadd_rn(double(j), a.ph[k][1]), yc);
        // Horner in v is shared by the whole row
        double r[2][3];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const double *c = cf + 2 + 9 * q;
            r[q][0] = dma(dma(c[2], v, c[1]), v, c[0]);
            r[q][1] = dma(dma(c[5], v, c[4]), v, c[3]);
            r[q][2] = dma(dma(c[8], v, c[7]), v, c[6]);
        }
        float *orow = out + ((long long)k * ny + j) * nx;
      for (int seg = 0; seg < segs; ++seg) {
        const int i = seg * 32 + lane;
        if (i >= nx) continue;
        const double u = __dsub_rn(__dadd_rn(double(i), phx), xc);
        double xy[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) xy[q] = dma(dma(r[q][2], u, r[q][1]), u, r[q][0]);
        const float x = __double2float_rn(xy[0]), y = __double2float_rn(xy[1]);
        float val = a.misval;
        if (x >= 0.f && x <= xmax && y >= 0.f && y <= ymax) {
            float t;
            if (INTERP == SPX_BLOT_POLY5)
                t = PREP == 2 ? poly5_prepared<true>(ws, w, x, y)
                  : PREP == 1 ? poly5_prepared<false>(ws, w, x, y) : poly5(d, h, w, x, y);
            else if (INTERP == SPX_BLOT_POLY3) t = poly3(d, h, w, x, y);
            else if (INTERP == SPX_BLOT_LINEAR) t = bilinear(d, h, w, x, y);
            else t = nearest(d, h, w, x, y);
            val = fm(t, scale);
        }
        orow[i] = val;
      }
    }
}

}  // namespace spx

extern "C" size_t spx_blot_workspace_bytes(int src_h, int src_w) {
    if (src_h < 1 || src_w < 1) return 0;
    return size_t(spx::prep_elems(src_h, src_w)) * sizeof(float4);
}

extern "C" int spx_blot_phases(const spx_blot_args *args_host, void *stream) {
    using namespace spx;
    if (!args_host) return SPX_E_BADARG;
    const spx_blot_args &g = *args_host;
    if (g.n < 0 || g.nph < 1 || g.nph > kMaxPhases || !g.phases_host) return SPX_E_BADARG;
    if (g.n == 0) return 0;
    if (!g.src || !g.src_offset || !g.src_shape || !g.map || !g.tgt_shape || !g.out_offset ||
        !g.out || g.max_ny < 1 || g.max_nx < 1)
        return SPX_E_BADARG;
    if (g.interp != SPX_BLOT_NEAREST && g.interp != SPX_BLOT_LINEAR && g.interp != SPX_BLOT_POLY5 &&
        g.interp != SPX_BLOT_POLY3)
        return SPX_E_UNSUPPORTED;
    if (g.n > 65535 * 64) return SPX_E_UNSUPPORTED;
    BlotDev a;
    a.src = g.src; a.src_off = reinterpret_cast<const long long *>(g.src_offset);
    a.src_shape = g.src_shape; a.map = g.map; a.scale = g.scale;
    a.tgt_shape = g.tgt_shape; a.out_off = reinterpret_cast<const long long *>(g.out_offset);
    a.out = g.out; a.nph = g.nph; a.interp = g.interp; a.misval = g.misval;
    const bool prep = g.interp == SPX_BLOT_POLY5 && g.workspace != nullptr;
    if (prep && (!g.ws_offset || g.max_src_h < 1 || g.max_src_w < 1)) return SPX_E_BADARG;
    // sources whose prepared table fits half an SM's shared memory: the table is built by
    // the CTA that uses it (chosen when the caller asked for the prepared path at all)
    const size_t smem_bytes = prep ? size_t(prep_elems(g.max_src_h, g.max_src_w)) * sizeof(float4) : 0;
    const bool smem_tab = prep && smem_bytes <= SPX_BLOT_SMEM_TABLE_MAX;
    if (smem_tab) {
        cudaError_t ea = cudaFuncSetAttribute(blot_kernel<SPX_BLOT_POLY5, 2>,
                                              cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem_bytes));
        if (ea != cudaSuccess) return int(ea);
    }
    a.ws = static_cast<float4 *>(g.workspace);
    a.ws_off = reinterpret_cast<const long long *>(g.ws_offset);
    for (int k = 0; k < g.nph; ++k) { a.ph[k][0] = g.phases_host[2 * k]; a.ph[k][1] = g.phases_host[2 * k + 1]; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // CTAs per source: a warp handles whole rows (phase, j), about 32 segments of 32 px each
    const long long items = (long long)g.nph * g.max_ny * ((g.max_nx + 31) / 32);
    int bx = int((items + 32 * (kBlotThreads / 32) - 1) / (32 * (kBlotThreads / 32)));
    if (bx < 1) bx = 1;
    if (bx > 64) bx = 64;
    for (int s0 = 0; s0 < g.n; s0 += 65535) {               // gridDim.y limit
        const int cnt = g.n - s0 < 65535 ? g.n - s0 : 65535;
        BlotDev b = a;
        b.n = cnt;
        b.src_off += s0; b.src_shape += 2 * s0; b.map += 20 * (long long)s0;
        if (b.scale) b.scale += s0;
        b.tgt_shape += 2 * s0; b.out_off += s0;
        if (prep) b.ws_off += s0;
        const dim3 grid(bx, cnt);
        if (smem_tab) {
            blot_kernel<SPX_BLOT_POLY5, 2><<<dim3(1, cnt), kBlotThreads, smem_bytes, st>>>(b);
        } else if (prep) {
            const long long pe = prep_elems(g.max_src_h, g.max_src_w);
            int px = int((pe + 4 * kBlotThreads - 1) / (4 * kBlotThreads));
            px = px < 1 ? 1 : (px > 64 ? 64 : px);
            blot_prep_kernel<<<dim3(px, cnt), kBlotThreads, 0, st>>>(b);
            blot_kernel<SPX_BLOT_POLY5, 1><<<grid, kBlotThreads, 0, st>>>(b);
        } else if (g.interp == SPX_BLOT_POLY5) blot_kernel<SPX_BLOT_POLY5, 0><<<grid, kBlotThreads, 0, st>>>(b);
        else if (g.interp == SPX_BLOT_POLY3) blot_kernel<SPX_BLOT_POLY3, 0><<<grid, kBlotThreads, 0, st>>>(b);
        else if (g.interp == SPX_BLOT_LINEAR) blot_kernel<SPX_BLOT_LINEAR, 0><<<grid, kBlotThreads, 0, st>>>(b);
        else blot_kernel<SPX_BLOT_NEAREST, 0><<<grid, kBlotThreads, 0, st>>>(b);
    }
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : int(e);
}
