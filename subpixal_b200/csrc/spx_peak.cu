// (4) Batched centroid.find_peak (reference centroid.py:18-236): one CTA per
// surface does the brute-force argmax (whole image, or the search box around a
// guess), one thread then runs the float64 quadratic fit of spx_peak.cuh.
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/subpixal_b200.h"
#include "spx_peak.cuh"

namespace spx {

constexpr int kPeakThreads = 128;

struct PeakParams {
    const void *surf; const unsigned char *mask; const double *guess;
    double *xy; int *peak; int *status;
    int n, h, w, swx, swy, fwx, fwy;
};

struct BestD { double v; int idx; };
__device__ __forceinline__ bool bestd_better(const BestD &a, const BestD &b) {
    const bool an = a.v != a.v, bn = b.v != b.v;
    if (an || bn) return an && (!bn || a.idx < b.idx);
    return a.v > b.v || (a.v == b.v && a.idx < b.idx);
}

__device__ __forceinline__ double py2round_dev(double x) {        // utils.py:144-161
    return x >= 0.0 ? floor(x + 0.5) : ceil(x - 0.5);
}

template <class T>
__global__ void __launch_bounds__(kPeakThreads, 8) find_peak_kernel(const PeakParams p) {
    __shared__ double s_v[kPeakThreads / 32];
    __shared__ int s_i[kPeakThreads / 32];
    const int k = blockIdx.x;
    const size_t npx = size_t(p.h) * p.w;
    const T *img = static_cast<const T *>(p.surf) + size_t(k) * npx;
    const unsigned char *msk = p.mask ? p.mask + size_t(k) * npx : nullptr;
    const int tid = threadIdx.x;

    // ---- where to look for the pre-fit maximum (centroid.py:111-156) ----
    int sx1 = 0, sx2 = p.w, sy1 = 0, sy2 = p.h;
    bool search = true, use_mask = msk != nullptr;
    int gi = 0, gj = 0;
    double gx = 0.0, gy = 0.0;
    if (p.guess) {
        gx = p.guess[2 * k]; gy = p.guess[2 * k + 1];
        gi = int(py2round_dev(gx)); gj = int(py2round_dev(gy));    // :130-131
        use_mask = false;                                          // :144-148 ignores mask
        if (p.swx > 0 && p.swy > 0) {                              // :134-152
            sx1 = max(0, gi - p.swx / 2); sx2 = min(p.w, sx1 + p.swx);
            sy1 = max(0, gj - p.swy / 2); sy2 = min(p.h, sy1 + p.swy);
            search = sx1 < sx2 && sy1 < sy2;
        } else {
            search = false;
        }
    }
    BestD best; best.v = -INFINITY; best.idx = 0x7fffffff;
    bool found = false;
    if (search) {
        const int bw = sx2 - sx1, cnt = bw * (sy2 - sy1);
        if (bw == p.w && sizeof(T) == 4 && !use_mask && (cnt & 3) == 0 &&
            (reinterpret_cast<size_t>(img + sy1 * p.w) & 15) == 0) {
            // contiguous float32 samples, 16-byte aligned: four 128-bit loads in flight per
            // thread.  float -> double is exact and monotonic, so the comparison runs in
            // float32 (the reference upcasts first, centroid.py:98): one NaN-propagating
            // max over the 16 samples decides whether the index search is needed at all.
            const int base = sy1 * p.w;
            const float4 *v4 = reinterpret_cast<const float4 *>(img + base);
            const int n4 = cnt >> 2;
            float bv = -INFINITY;           // running first maximum of this thread (NaN wins)
            int bi = 0x7fffffff;
            for (int t0 = tid; t0 < n4; t0 += 4 * kPeakThreads) {
                float4 q[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int t = t0 + u * kPeakThreads;
                    q[u] = t < n4 ? __ldg(v4 + t) : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
                }
                float m = -INFINITY;
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    float m4;
                    asm("max.NaN.f32 %0, %1, %2;" : "=f"(m4) : "f"(q[u].x), "f"(q[u].y));
                    asm("max.NaN.f32 %0, %0, %1;" : "+f"(m4) : "f"(q[u].z));
                    asm("max.NaN.f32 %0, %0, %1;" : "+f"(m4) : "f"(q[u].w));
                    asm("max.NaN.f32 %0, %0, %1;" : "+f"(m) : "f"(m4));
                }
                // padding (-inf) never wins: bi stays unset until a real sample is seen
                if ((bv == bv) && (!(m <= bv) || bi == 0x7fffffff)) {
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int t = t0 + u * kPeakThreads;
                        if (t >= n4) break;
                        const float va[4] = {q[u].x, q[u].y, q[u].z, q[u].w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float v = va[e];
                            if ((bv == bv) && (v != v || v > bv || bi == 0x7fffffff)) {
                                bv = v; bi = base + 4 * t + e;
                            }
                        }
                    }
                }
            }
            if (bi != 0x7fffffff) { best.v = double(bv); best.idx = bi; found = true; }
        } else if (bw == p.w) {
            // full-width window: the samples are contiguous, no div/mod, four
            // independent loads in flight per thread
            const int base = sy1 * p.w;
            for (int t0 = tid; t0 < cnt; t0 += 4 * kPeakThreads) {
                T v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int t = t0 + u * kPeakThreads;
                    v[u] = t < cnt ? img[base + t] : T(0);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int t = t0 + u * kPeakThreads;
                    if (t >= cnt) break;
                    const int idx = base + t;
                    if (use_mask && !msk[idx]) continue;
                    BestD c; c.v = double(v[u]); c.idx = idx;
                    if (!found || bestd_better(c, best)) { best = c; found = true; }
                }
            }
        } else {
            for (int t = tid; t < cnt; t += kPeakThreads) {
                const int j = sy1 + t / bw, i = sx1 + t % bw;
                const int idx = j * p.w + i;
                if (use_mask && !msk[idx]) continue;
                BestD c; c.v = double(img[idx]); c.idx = idx;
                if (!found || bestd_better(c, best)) { best = c; found = true; }
            }
        }
        if (!found) { best.v = -INFINITY; best.idx = 0x7fffffff; }
        // NaN must be able to win over "nothing found": encode absence by idx
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            BestD t;
            t.v = __shfl_xor_sync(0xffffffffu, best.v, o);
            t.idx = __shfl_xor_sync(0xffffffffu, best.idx, o);
            if (t.idx != 0x7fffffff && (best.idx == 0x7fffffff || bestd_better(t, best))) best = t;
        }
        if ((tid & 31) == 0) { s_v[tid >> 5] = best.v; s_i[tid >> 5] = best.idx; }
    }
    __syncthreads();
    if (tid != 0) return;

    int imax = gi, jmax = gj;
    double cx = gx, cy = gy;
    int st = 0;
    if (search) {
        BestD g; g.v = -INFINITY; g.idx = 0x7fffffff;
        for (int w = 0; w < kPeakThreads / 32; ++w) {
            BestD t; t.v = s_v[w]; t.idx = s_i[w];
            if (t.idx != 0x7fffffff && (g.idx == 0x7fffffff || bestd_better(t, g))) g = t;
        }
        if (g.idx == 0x7fffffff) {        // empty mask: numpy raises; report it
            p.xy[2 * k] = 0.0; p.xy[2 * k + 1] = 0.0; p.peak[k] = -1;
            p.status[k] = SPX_S_FEW_POINTS;
            return;
        }
        jmax = g.idx / p.w; imax = g.idx - jmax * p.w;
        cx = double(imax); cy = double(jmax);
    }
    PeakOut po = fit_peak([&](int j, int i) { return double(img[size_t(j) * p.w + i]); },
                          [&](int j, int i) { return msk ? msk[size_t(j) * p.w + i] != 0 : true; },
                          p.h, p.w, imax, jmax, cx, cy, p.fwx, p.fwy);
    st |= po.status;
    p.xy[2 * k] = po.x; p.xy[2 * k + 1] = po.y;
    p.peak[k] = (imax >= 0 && imax < p.w && jmax >= 0 && jmax < p.h) ? jmax * p.w + imax : -1;
    p.status[k] = st;
}

}  // namespace spx

extern "C" int spx_find_peak(const void *surfaces, int kind, int n, int h, int w,
                             const uint8_t *mask, const double *guess,
                             int search_wx, int search_wy, int fit_wx, int fit_wy,
                             double *xy, int32_t *peak, int32_t *status, void *stream) {
    using namespace spx;
    if (!surfaces || !xy || !peak || !status || n < 0 || h < 1 || w < 1) return SPX_E_BADARG;
    if (fit_wx < 1 || fit_wy < 1 || search_wx < 0 || search_wy < 0) return SPX_E_BADARG;
    if (fit_wx > SPX_MAX_FIT_BOX || fit_wy > SPX_MAX_FIT_BOX) return SPX_E_UNSUPPORTED;
    if (kind != SPX_T_F32 && kind != SPX_T_F64) return SPX_E_DTYPE;
    if (size_t(h) * size_t(w) > size_t(0x7ffffffe)) return SPX_E_UNSUPPORTED;
    if (n == 0) return 0;
    PeakParams p{};
    p.surf = surfaces; p.mask = mask; p.guess = guess; p.xy = xy; p.peak = peak; p.status = status;
    p.n = n; p.h = h; p.w = w; p.swx = search_wx; p.swy = search_wy; p.fwx = fit_wx; p.fwy = fit_wy;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (kind == SPX_T_F32) find_peak_kernel<float><<<n, kPeakThreads, 0, st>>>(p);
    else find_peak_kernel<double><<<n, kPeakThreads, 0, st>>>(p);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : int(e);
}
