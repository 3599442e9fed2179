// (1) Batched cutout gather: the crop / fill / mask arithmetic of
// cutout.Cutout.__init__ (reference cutout.py:737-779) plus the mask updates
// of the list builders (cutout.py:188-196, 286-292, 416-427) and the optional
// masked-pixel zeroing of align.py:661, for n boxes in one launch.
//
// HBM-bound byte work: one warp per cutout row, lanes along x, so every
// global load/store instruction touches one contiguous row segment.
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/subpixal_b200.h"
#include "spx_common.cuh"

namespace spx {

constexpr int kGatherThreads = 256;

struct GatherDev {
    const void *frame; const void *dq; const void *weight; const void *seg;
    const long long *src_ids; const int *boxes; const long long *out_offset;
    void *data_out; unsigned char *mask_out; void *dq_out; void *weight_out;
    long long frame_pitch, dq_good_bits;
    double fillval;
    int H, W, n, dq_bytes, weight_f64, seg_bytes, dq_use_bits, zero_masked;
};

__device__ __forceinline__ long long load_int(const void *p, long long i, int bytes) {
    switch (bytes) {
        case 1: return static_cast<const int8_t *>(p)[i];
        case 2: return static_cast<const int16_t *>(p)[i];
        case 4: return static_cast<const int32_t *>(p)[i];
        default: return static_cast<const long long *>(p)[i];
    }
}
// DQ words are bit fields: zero-extend so that the bits above the element
// width never take part in the (dq & ~good) test
__device__ __forceinline__ long long load_bits(const void *p, long long i, int bytes) {
    switch (bytes) {
        case 1: return static_cast<const uint8_t *>(p)[i];
        case 2: return static_cast<const uint16_t *>(p)[i];
        case 4: return static_cast<const uint32_t *>(p)[i];
        default: return static_cast<const long long *>(p)[i];
    }
}
__device__ __forceinline__ void store_int(void *p, long long i, int bytes, long long v) {
    switch (bytes) {
        case 1: static_cast<int8_t *>(p)[i] = int8_t(v); break;
        case 2: static_cast<int16_t *>(p)[i] = int16_t(v); break;
        case 4: static_cast<int32_t *>(p)[i] = int32_t(v); break;
        default: static_cast<long long *>(p)[i] = v; break;
    }
}
template <class T> __device__ __forceinline__ bool is_finite(T v);
template <> __device__ __forceinline__ bool is_finite<float>(float v) { return finitef(v); }
template <> __device__ __forceinline__ bool is_finite<double>(double v) {
    return (__double2hiint(v) & 0x7ff00000) != 0x7ff00000;
}

// Each warp owns kRowsPerWarp rows of the cutout and walks them together, so a
// lane has kRowsPerWarp independent global loads in flight per step (the kernel
// is latency bound otherwise: one 128-byte request per warp is a fraction of
// what Little's law needs to saturate HBM).  AUX = dq / segmentation / weight
// planes present (their loads are batched with the pixel loads).
#ifndef SPX_GATHER_RPW
#define SPX_GATHER_RPW 4
#endif
constexpr int kRowsPerWarp = SPX_GATHER_RPW;
constexpr int kGatherRowsPerCta = kRowsPerWarp * (kGatherThreads / 32);

// DQB / SEGB: byte width of the dq / segmentation elements known at compile time
// (0 = plane absent, -1 = decide at run time).  The common combinations (uint16 dq,
// int32 / int64 segmentation) get their own instantiations so that the auxiliary
// loads are plain, batched loads instead of a per-pixel switch on the width.
template <int B> __device__ __forceinline__ long long load_bits_c(const void *p, long long i, int bytes) {
    if constexpr (B == 2) return static_cast<const uint16_t *>(p)[i];
    else if constexpr (B == 4) return static_cast<const uint32_t *>(p)[i];
    else return load_bits(p, i, bytes);
}
template <int B> __device__ __forceinline__ long long load_int_c(const void *p, long long i, int bytes) {
    if constexpr (B == 4) return static_cast<const int32_t *>(p)[i];
    else if constexpr (B == 8) return static_cast<const long long *>(p)[i];
    else return load_int(p, i, bytes);
}

template <class T, bool AUX, int DQB = -1, int SEGB = -1>
__global__ void __launch_bounds__(kGatherThreads) gather_kernel(const GatherDev a) {
    const int k = blockIdx.y;
    const int4 box = __ldg(reinterpret_cast<const int4 *>(a.boxes) + k);
    const int x1 = box.x, y1 = box.y, w = box.z - box.x + 1, h = box.w - box.y + 1;
    const int row0 = blockIdx.x * kGatherRowsPerCta;
    if (row0 >= h || w <= 0) return;
    const long long off = a.out_offset[k];
    const long long sid = (AUX && a.src_ids) ? a.src_ids[k] : 0;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const T *frame = static_cast<const T *>(a.frame);
    T *out = static_cast<T *>(a.data_out);
    const T fill = T(a.fillval);
    const bool want_dq = AUX && DQB != 0 && a.dq != nullptr;
    const bool want_seg = AUX && SEGB != 0 && a.seg && a.src_ids;
    const bool want_w = AUX && a.weight && a.weight_out;
    const int rbase = row0 + warp * kRowsPerWarp;
    for (int col = lane; col < w; col += 32) {
        const int xx = x1 + col;
        const bool incol = xx >= 0 && xx < a.W;
        T v[kRowsPerWarp];
        long long dqv[kRowsPerWarp], sgv[kRowsPerWarp];
        bool inside[kRowsPerWarp];
#pragma unroll
        for (int r = 0; r < kRowsPerWarp; ++r) {                       // pixel loads first
            const int row = rbase + r, yy = y1 + row;
            inside[r] = incol && row < h && yy >= 0 && yy < a.H;      // cutout.py:742-752
            v[r] = inside[r] ? __ldg(frame + (long long)yy * a.frame_pitch + xx) : fill;
        }
        if (AUX) {                                                     // then the auxiliary planes, batched
#pragma unroll
            for (int r = 0; r < kRowsPerWarp; ++r) {
                const long long src = (long long)(y1 + rbase + r) * a.W + xx;
                dqv[r] = (want_dq && inside[r]) ? load_bits_c<DQB>(a.dq, src, a.dq_bytes) : 0;
            }
#pragma unroll
            for (int r = 0; r < kRowsPerWarp; ++r) {
                const long long src = (long long)(y1 + rbase + r) * a.W + xx;
                sgv[r] = (want_seg && inside[r]) ? load_int_c<SEGB>(a.seg, src, a.seg_bytes) : 0;
            }
        }
#pragma unroll
        for (int r = 0; r < kRowsPerWarp; ++r) {
            const int row = rbase + r;
            if (row >= h) break;
            const long long o = off + (long long)row * w + col;
            T val = v[r];
            bool bad = !inside[r] || !is_finite<T>(val);               // cutout.py:755
            if (AUX) {
                const long long src = (long long)(y1 + row) * a.W + xx;
                if (want_dq) {                                         // cutout.py:758-766
                    if (a.dq_out) store_int(a.dq_out, o, a.dq_bytes, dqv[r]);
                    if (a.dq_use_bits) bad |= (dqv[r] & ~a.dq_good_bits) != 0;   // cutout.py:190-193
                }
                if (want_seg) bad |= (sgv[r] != sid);                  // cutout.py:188, 416-419
                if (want_w) {                                          // cutout.py:769-779
                    if (a.weight_f64)
                        static_cast<double *>(a.weight_out)[o] =
                            inside[r] ? static_cast<const double *>(a.weight)[src] : 0.0;
                    else
                        static_cast<float *>(a.weight_out)[o] =
                            inside[r] ? static_cast<const float *>(a.weight)[src] : 0.f;
                }
            }
            if (a.zero_masked && bad) val = T(0);                      // align.py:661
            out[o] = val;
            if (a.mask_out) a.mask_out[o] = bad ? 1 : 0;
        }
    }
}

static int int_bytes(int kind) {
    switch (kind) {
        case SPX_T_I8: return 1; case SPX_T_I16: return 2;
        case SPX_T_I32: return 4; case SPX_T_I64: return 8; default: return 0;
    }
}

}  // namespace spx

extern "C" int spx_gather_cutouts(const spx_gather_args *g, void *stream) {
    using namespace spx;
    if (!g || !g->frame || !g->boxes || !g->out_offset || !g->data_out) return SPX_E_BADARG;
    if (g->n < 0 || g->frame_h < 1 || g->frame_w < 1 || g->frame_pitch < g->frame_w ||
        g->max_h < 1 || g->max_w < 1) return SPX_E_BADARG;
    if (g->frame_kind != SPX_T_F32 && g->frame_kind != SPX_T_F64) return SPX_E_DTYPE;
    GatherDev a{};
    a.frame = g->frame; a.dq = g->dq; a.weight = g->weight; a.seg = g->seg;
    a.src_ids = reinterpret_cast<const long long *>(g->src_ids);
    a.boxes = g->boxes; a.out_offset = reinterpret_cast<const long long *>(g->out_offset);
    a.data_out = g->data_out; a.mask_out = g->mask_out; a.dq_out = g->dq_out;
    a.weight_out = g->weight_out;
    a.frame_pitch = g->frame_pitch; a.dq_good_bits = g->dq_good_bits;
    a.fillval = g->fillval; a.H = g->frame_h; a.W = g->frame_w; a.n = g->n;
    a.dq_use_bits = g->dq_use_bits; a.zero_masked = g->zero_masked;
    if (g->dq) { a.dq_bytes = int_bytes(g->dq_kind); if (!a.dq_bytes) return SPX_E_DTYPE; }
    if (g->seg) { a.seg_bytes = int_bytes(g->seg_kind); if (a.seg_bytes < 4) return SPX_E_DTYPE; }
    if (g->weight) {
        if (g->weight_kind != SPX_T_F32 && g->weight_kind != SPX_T_F64) return SPX_E_DTYPE;
        a.weight_f64 = g->weight_kind == SPX_T_F64;
    }
    if (g->n == 0) return 0;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int rows = (g->max_h + kGatherRowsPerCta - 1) / kGatherRowsPerCta;
    const bool aux = g->dq || (g->seg && g->src_ids) || (g->weight && g->weight_out);
    for (int k0 = 0; k0 < g->n; k0 += 65535) {              // gridDim.y limit
        const int nk = (g->n - k0 < 65535) ? g->n - k0 : 65535;
        GatherDev b = a;
        b.boxes = a.boxes + 4 * size_t(k0);
        b.out_offset = a.out_offset + k0;
        if (a.src_ids) b.src_ids = a.src_ids + k0;
        dim3 grid(rows, nk);
        if (g->frame_kind == SPX_T_F32) {
            const bool has_seg = g->seg && g->src_ids;
            const int dqb = g->dq ? a.dq_bytes : 0, sgb = has_seg ? a.seg_bytes : 0;
            if (!aux) gather_kernel<float, false><<<grid, kGatherThreads, 0, st>>>(b);
            else if (dqb == 2 && sgb == 4) gather_kernel<float, true, 2, 4><<<grid, kGatherThreads, 0, st>>>(b);
            else if (dqb == 2 && sgb == 8) gather_kernel<float, true, 2, 8><<<grid, kGatherThreads, 0, st>>>(b);
            else if (dqb == 2 && sgb == 0) gather_kernel<float, true, 2, 0><<<grid, kGatherThreads, 0, st>>>(b);
            else if (dqb == 0 && sgb == 4) gather_kernel<float, true, 0, 4><<<grid, kGatherThreads, 0, st>>>(b);
            else if (dqb == 0 && sgb == 8) gather_kernel<float, true, 0, 8><<<grid, kGatherThreads, 0, st>>>(b);
            else if (dqb == 4 && sgb == 4) gather_kernel<float, true, 4, 4><<<grid, kGatherThreads, 0, st>>>(b);
            else gather_kernel<float, true><<<grid, kGatherThreads, 0, st>>>(b);
        } else {
            if (aux) gather_kernel<double, true><<<grid, kGatherThreads, 0, st>>>(b);
            else gather_kernel<double, false><<<grid, kGatherThreads, 0, st>>>(b);
        }
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return int(e);
    }
    return 0;
}
