// (5) Bounding boxes of every label of a segmentation image in ONE pass over the
// image (the reference scans the whole frame once per source: `seg == sid`,
// np.where, min/max - cutout.py:151-162, O(n_src * H * W)).
// Integer work, HBM-bound: each warp reads 32 consecutive pixels of a row,
// lanes holding the same label elect a leader (match_any) which issues four
// atomics for the whole run.
#include <cuda_runtime.h>
#include <limits.h>
#include <stdint.h>
#include "../../include/subpixal_b200.h"

namespace spx {

__global__ void segbox_init_kernel(int *boxes, int nlab) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nlab) {
        boxes[4 * i + 0] = INT_MAX; boxes[4 * i + 1] = -1;     // x1, x2
        boxes[4 * i + 2] = INT_MAX; boxes[4 * i + 3] = -1;     // y1, y2
    }
}

template <class T>
__global__ void __launch_bounds__(256) segbox_kernel(const T *seg, int h, int w, long long pitch,
                                                       int max_label, int *boxes) {
    const int lane = threadIdx.x & 31;
    const int warps_per_row = (w + 31) / 32;
    const long long nwarp_items = (long long)h * warps_per_row;
    const long long wid0 = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long wstride = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long it = wid0; it < nwarp_items; it += wstride) {
        const int y = int(it / warps_per_row);
        const int x = int(it - (long long)y * warps_per_row) * 32 + lane;
        long long lab = 0;
        if (x < w) lab = (long long)seg[(long long)y * pitch + x];
        if (lab < 0 || lab > max_label) lab = 0;
        const unsigned same = __match_any_sync(0xffffffffu, int(lab));
        if (lab > 0 && lane == __ffs(same) - 1) {
            const int xlo = x, xhi = x - lane + (31 - __clz(same));
            int *b = boxes + 4 * lab;
            atomicMin(b + 0, xlo); atomicMax(b + 1, xhi);
            atomicMin(b + 2, y);   atomicMax(b + 3, y);
        }
    }
}

}  // namespace spx

extern "C" int spx_segment_boxes(const void *seg, int seg_kind, int h, int w, int64_t pitch,
                                 int max_label, int32_t *boxes_out, void *stream) {
    using namespace spx;
    if (!seg || !boxes_out || h < 1 || w < 1 || pitch < w || max_label < 0) return SPX_E_BADARG;
    if (seg_kind != SPX_T_I32 && seg_kind != SPX_T_I64) return SPX_E_DTYPE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int nlab = max_label + 1;
    segbox_init_kernel<<<(nlab + 255) / 256, 256, 0, st>>>(boxes_out, nlab);
    const long long items = (long long)h * ((w + 31) / 32);
    long long blocks = (items + 7) / 8;
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (seg_kind == SPX_T_I32)
        segbox_kernel<int32_t><<<int(blocks), 256, 0, st>>>(static_cast<const int32_t *>(seg), h, w,
                                                             pitch, max_label, boxes_out);
    else
        segbox_kernel<long long><<<int(blocks), 256, 0, st>>>(static_cast<const long long *>(seg), h,
                                                               w, pitch, max_label, boxes_out);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : int(e);
}
