/*
 * subpixal_b200 - C ABI of the B200 (sm_100a) implementation of subpixal's
 * data-parallel hot path:  cutout gather -> 4-phase cross-correlation ->
 * interlace -> sub-pixel peak fit.
 *
 * The reference (spacetelescope/subpixal) is pure Python and has no FFI of its
 * own; its boundary for this path is the Python API of three modules
 * (SURVEY.md section 8b).  Each entry point below states which reference
 * function(s) it replaces.  The Python package `subpixal_b200` binds these
 * symbols with ctypes and keeps the reference's signatures on top of them
 * (see INTEGRATION.md for the stub a subpixal maintainer would add).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer on the current CUDA device unless the
 *     name ends in `_host`; the library never allocates, never synchronises
 *     the device and never throws: work is enqueued on `stream`
 *     (a cudaStream_t / CUstream passed as void*; NULL = legacy default).
 *   - return value: 0 = ok, < 0 = SPX_E_* (bad argument / unsupported size),
 *     > 0 = a cudaError_t raised while launching.
 *   - images are row-major, `pitch` counted in ELEMENTS.
 *   - there is no CPU fallback anywhere in this library.
 */
#ifndef SUBPIXAL_B200_H
#define SUBPIXAL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPX_ABI_VERSION 1

/* ---- error codes (negative return values) ------------------------------ */
#define SPX_E_BADARG        (-1)  /* NULL pointer, negative size, bad enum   */
#define SPX_E_UNSUPPORTED   (-2)  /* cutout larger than SPX_MAX_CUTOUT, ...  */
#define SPX_E_WORKSPACE     (-3)  /* workspace missing or too small          */
#define SPX_E_DTYPE         (-4)  /* unsupported element type                */

#define SPX_MAX_CUTOUT 256        /* max ny, nx handled by the CC kernels    */
#define SPX_MIN_CUTOUT 3          /* 2n >= 6 so the 5x5 fit box exists       */

/* ---- cc_type (cc.py:107-111; anything else behaves as CC, cc.py:108) ---- */
#define SPX_CC   0
#define SPX_NCC  1
#define SPX_ZNCC 2

/* ---- flags -------------------------------------------------------------- */
#define SPX_F_ZERO_BAD_PHASES 0x1u /* fused path: non-finite pixels of the four
                                      phase frames are treated as masked and
                                      set to 0 (align.py:661 + cutout.py:755) */
#define SPX_F_ZERO_BAD_REF    0x2u /* same for the image ("ref_image") frame;
                                      NOT reference behaviour, off by default */
#define SPX_F_EXACT_NORM      0x4u /* NCC without a requested surface normally skips the
                                      statistics sweep when the standard deviations of
                                      cc.py:139-154 are provably positive (the shift does
                                      not depend on a positive scale); this flag forces
                                      the sweep */

/* ---- per-pair status bits (int32 output) --------------------------------
 * The reference expresses these situations through NaNs and through the
 * fall-back return values of centroid.find_peak; the bits let the Python
 * wrapper reproduce them exactly. */
#define SPX_S_NONFINITE_REF   0x001 /* NaN/Inf in ref_image -> ICC all NaN    */
#define SPX_S_NONFINITE_P00   0x002 /* NaN/Inf in image00 -> that lattice NaN */
#define SPX_S_NONFINITE_P10   0x004
#define SPX_S_NONFINITE_P01   0x008
#define SPX_S_NONFINITE_P11   0x010
#define SPX_S_BAD_NORM        0x020 /* std is 0/NaN (cc.py:139-142) -> NaN ICC */
#define SPX_S_PEAK_EDGE       0x040 /* centroid.py:171-172 integer position   */
#define SPX_S_NO_MAXIMUM      0x080 /* centroid.py:217-225 box centre         */
#define SPX_S_FIT_OUTSIDE     0x100 /* centroid.py:230-236 integer position   */
#define SPX_S_FIT_NONFINITE   0x200 /* NaN in fit box -> integer position     */
#define SPX_S_BOX_MISMATCH    0x400 /* fused path: box size != (ny, nx)       */
#define SPX_S_FEW_POINTS      0x800 /* < 6 usable points (centroid.py:160,186,202) */
#define SPX_S_SINGULAR       0x1000 /* rank-deficient masked fit (find_peak only) */
#define SPX_S_UNSUPPORTED_SIZE 0x2000 /* ragged batch: an axis < 3 or > 256 px; shift NaN,
                                        peak -1 (set by the host binding, the pair is not launched) */

/* ---- element kinds for the gather --------------------------------------- */
#define SPX_T_F32 0
#define SPX_T_F64 1
#define SPX_T_I8  2   /* integers: only the byte width matters */
#define SPX_T_I16 3
#define SPX_T_I32 4
#define SPX_T_I64 5

int spx_abi_version(void);
const char *spx_error_string(int code);

/* ========================================================================
 * (1) Batched cutout gather.
 * Replaces the crop/mask arithmetic of cutout.Cutout.__init__
 * (cutout.py:689-797) and the mask updates of the list builders
 * (cutout.py:188-196, 286-292, 416-427) and, with `zero_masked`, the
 * sentinel write of align.py:661 - for n boxes in one launch.
 *
 * boxes      [n,4] int32 (x1, y1, x2, y2), trc INCLUSIVE (cutout.py:1024-1031);
 *            boxes may stick out of the frame ('fill' mode); the caller
 *            rejects 'strict' violations / empty overlaps on the host.
 * out_offset [n] int64: element offset of cutout k in every output buffer
 *            (ragged packing; k*h*w for uniform boxes).
 * data_out   frame dtype; pixels outside the frame get `fillval`.
 * mask_out   uint8, 1 = bad: outside frame | non-finite | seg != src_id |
 *            (dq & ~dq_good_bits) != 0.
 * dq/weight/seg and their outputs may be NULL.  dq_use_bits = 0 skips the
 * DQ test (dqbitmask=None).  seg test is skipped when src_ids is NULL.
 * ======================================================================== */
typedef struct spx_gather_args {
    const void *frame;      int frame_kind;   /* SPX_T_F32 / SPX_T_F64       */
    int frame_h, frame_w;   int64_t frame_pitch;
    const void *dq;         int dq_kind;      /* SPX_T_I8..I64, same shape   */
    const void *weight;     int weight_kind;  /* SPX_T_F32 / SPX_T_F64       */
    const void *seg;        int seg_kind;     /* SPX_T_I32 / SPX_T_I64       */
    const int64_t *src_ids;                   /* [n] or NULL                 */
    int dq_use_bits;        int64_t dq_good_bits;
    const int32_t *boxes;   const int64_t *out_offset;   int n;
    int max_h, max_w;                         /* upper bounds of box sizes   */
    double fillval;         int zero_masked;
    void *data_out;  uint8_t *mask_out;  void *dq_out;  void *weight_out;
} spx_gather_args;

int spx_gather_cutouts(const spx_gather_args *args_host, void *stream);

/* ========================================================================
 * (2) Batched cross-correlation + sub-pixel fit on cutout stacks.
 * Replaces cc.find_displacement / _build_icc_image / _normalize
 * (cc.py:21-156) and the centroid.find_peak call at cc.py:86 for n
 * independent problems.
 *
 * ref, im00, im10, im01, im11: float32 [n][ny][nx] with the given strides
 *            (elements) between consecutive pairs - a [n,5,ny,nx] stack is
 *            ref=base, im00=base+ny*nx, ..., stride=5*ny*nx.
 * shifts     [n,2] float64 (dx, dy)      (cc.py:92-93)
 * peak       [n] int32 flat argmax index into the (2ny, 2nx) surface
 * status     [n] int32 SPX_S_* bits
 * icc        [n,2ny,2nx] float32 or NULL (full_output, cc.py:95,128)
 * workspace  spx_cc_workspace_bytes(...) bytes, 256-byte aligned
 *
 * Packed output (all three spx_cc_fit_* entry points): when `peak` and
 * `status` are BOTH NULL, `shifts` is an [n,4] float64 table and pair p gets
 * the row (dx, dy, peak, status) - the per-rank block that the multi-GPU path
 * sends with its single all-gather (align.py:689 collects exactly this table),
 * written by the kernel itself so that no packing kernels run per step.
 * ======================================================================== */
size_t spx_cc_workspace_bytes(int n, int ny, int nx, int want_icc);

int spx_cc_fit_stack(const float *ref, const float *im00, const float *im10,
                     const float *im01, const float *im11,
                     const int64_t pair_stride_host[5],
                     int n, int ny, int nx, int cc_type,
                     double *shifts, int32_t *peak, int32_t *status,
                     float *icc, void *workspace, size_t workspace_bytes,
                     void *stream);

/* (2b) The same for cutouts of DIFFERENT shapes (the bounding-box cutouts of
 * create_cutouts are ragged, cutout.py:159-175): pair p occupies the
 * contiguous block [5][ny_p][nx_p] (ref_image, image00, image10, image01,
 * image11) at element offsets[p] of `stack`, shapes[p] = (ny_p, nx_p).
 * order      [n] int32 device: the pairs this launch processes (outputs are
 *            written at index p = order[k], so several launches can fill one
 *            table), or NULL for 0..n-1.  Every processed pair must fit the
 *            size class of (max_ny, max_nx) - 20, 32, 42, 64, 84, 128, 170 or 256 px; the
 *            caller buckets pairs by class (one launch per class) because a
 *            pair costs as much as the largest shape of its class.
 * icc        ragged output: surface p at element icc_offsets[p], or NULL.
 * workspace  spx_cc_workspace_bytes(n, max_ny, max_nx, icc != NULL) bytes. */
int spx_cc_fit_ragged(const float *stack, const int64_t *offsets, const int32_t *shapes,
                      const int32_t *order, int n, int max_ny, int max_nx, int cc_type,
                      double *shifts, int32_t *peak, int32_t *status,
                      float *icc, const int64_t *icc_offsets,
                      void *workspace, size_t workspace_bytes, void *stream);

/* ========================================================================
 * (3) Fused gather + cross-correlation + fit straight from full frames.
 * = (1) with fillval 0 for five frames sharing one box table, feeding (2)
 * without the cutouts ever being written to HBM.  frames_host[0] is the
 * image frame (ref_image), [1..4] the reference frames rendered at the four
 * half-pixel phases (image00, image10, image01, image11).
 * boxes [n,4] int32 (x1,y1,x2,y2) must all be nx wide and ny high.
 * ======================================================================== */
int spx_cc_fit_frames(const float *const frames_host[5],
                      int frame_h, int frame_w, int64_t frame_pitch,
                      const int32_t *boxes, int n, int ny, int nx,
                      int cc_type, unsigned flags,
                      double *shifts, int32_t *peak, int32_t *status,
                      float *icc, void *workspace, size_t workspace_bytes,
                      void *stream);

/* ========================================================================
 * (4) Batched centroid.find_peak (centroid.py:18-236) on n surfaces.
 * surfaces   [n][h][w] float32 or float64 (kind), contiguous.
 * mask       [n][h][w] uint8, 1 = GOOD (find_peak's convention), or NULL.
 * guess      [n,2] float64 (xmax, ymax) or NULL (global argmax).
 * search_wx/wy: brute-force search box around the guess; 0 = no search
 *            (ignored when guess is NULL - centroid.py:111-127).
 * fit_wx/wy  peak_fit_box (<= SPX_MAX_FIT_BOX each).
 * xy         [n,2] float64; status [n] int32; peak [n] int32 flat index of
 *            the pre-fit pixel.
 * The auto_expand_search recursion of centroid.py:210-234 is left to the
 * caller: it re-invokes with guess=NULL for the surfaces whose status has
 * NO_MAXIMUM / FIT_OUTSIDE set.
 * ======================================================================== */
#define SPX_MAX_FIT_BOX 15

int spx_find_peak(const void *surfaces, int kind, int n, int h, int w,
                  const uint8_t *mask, const double *guess,
                  int search_wx, int search_wy, int fit_wx, int fit_wy,
                  double *xy, int32_t *peak, int32_t *status, void *stream);

/* ========================================================================
 * (5) Bounding boxes of all labels of a segmentation image in one pass.
 * Replaces the per-source full-frame scan of create_primary_cutouts
 * (cutout.py:151-162: `seg == sid`, np.where, min / max for every source).
 * seg        [h][w] int32 or int64 (SPX_T_I32 / SPX_T_I64), pitch in elements.
 * boxes_out  [max_label+1, 4] int32 (x1, x2, y1, y2) per label; labels that do
 *            not occur (and label 0, the background) get (INT_MAX,-1,INT_MAX,-1).
 *            Labels outside [1, max_label] are ignored.
 * ======================================================================== */
int spx_segment_boxes(const void *seg, int seg_kind, int h, int w, int64_t pitch,
                      int max_label, int32_t *boxes_out, void *stream);

/* ========================================================================
 * (6) Batched blot of source (drizzled) cutouts onto target (image) cutout
 * grids at several sub-pixel phases.
 * Replaces blot.blot_cutout (blot.py:79-155; drizzlepac cdriz.tblot with
 * interp='poly5', misval=0, kscale=1) as called four times per source by
 * align.find_linear_fit (align.py:661-679: target grid displaced by half a
 * pixel along x, along x and y, along y).
 *
 * src        ragged stack of float32 source cutouts (masked pixels already
 *            zeroed, align.py:661); cutout k starts at element src_offset[k]
 *            and has shape src_shape[k] = (h, w).
 * map        [n,20] float64: xc, yc, a[3][3], b[3][3] - source position
 *            (0-based pixels) of target pixel (x, y), u = x - xc, v = y - yc:
 *              xs = sum_pq a[p][q] u^p v^q,  ys = sum_pq b[p][q] u^p v^q
 *            (the host samples the WCS map of blot.py:21-76 on a 3x3 grid).
 * phases_host [nph,2] float64 (ddx, ddy): phase k evaluates the map at
 *            (i + ddx, j + ddy); nph <= 8.  align.py's four blots are
 *            (0,0), (0.5,0), (0,0.5), (0.5,0.5) (Cutout.dx -= 0.5 moves the
 *            grid by +0.5 px in the image frame, cutout.py:1112-1145).
 * scale      [n] float32 or NULL: flux factor (ef / pix_ratio^2 and the
 *            counts/rate conversions of blot.py:134-138,148-150).
 * tgt_shape  [n,2] int32 (ny, nx); out_offset [n] int64: element offset of
 *            phase 0 of cutout k in `out`; its phases follow contiguously
 *            ([nph][ny][nx]).  max_ny, max_nx: largest target (grid sizing).
 * Pixels that map outside [0, w-1] x [0, h-1] of their source get `misval`.
 * workspace  optional (poly5 only): 16-byte aligned scratch in which the sources
 *            are prepared once (edge-extended sample + its x-direction central
 *            differences per pixel) so that the four phases share that work;
 *            source k uses spx_blot_workspace_bytes(h_k, w_k) bytes starting at
 *            16 * ws_offset[k] (ws_offset [n] int64, in 16-byte units);
 *            max_src_h/max_src_w: largest source.  NULL = direct evaluation.
 *            Results are bit-identical either way.
 * ======================================================================== */
#define SPX_BLOT_NEAREST 0
#define SPX_BLOT_LINEAR  1
#define SPX_BLOT_POLY5   2
#define SPX_BLOT_POLY3   3   /* 4x4 cubic (drizzle's 'poly3') */

typedef struct spx_blot_args {
    const float *src; const int64_t *src_offset; const int32_t *src_shape;
    const double *map; const float *scale;
    const int32_t *tgt_shape; const int64_t *out_offset;
    float *out;
    const double *phases_host;
    int n, nph, interp, max_ny, max_nx;
    float misval;
    void *workspace; const int64_t *ws_offset; int max_src_h, max_src_w;
} spx_blot_args;

size_t spx_blot_workspace_bytes(int src_h, int src_w);
int spx_blot_phases(const spx_blot_args *args_host, void *stream);

/* ========================================================================
 * (7) FITS image payload -> native device array (section 8f rank 4: ingest).
 * Replaces what astropy.io.fits does when align.py reads `h[ext].data`
 * (align.py:311-320, 466-486) and catalogs.py:757 `fits.getdata`: big-endian
 * two's-complement / IEEE-754 elements (FITS standard 4.0, sections 5.2, 5.3)
 * to native order, the BSCALE / BZERO linear scaling (section 4.4.2.5), and -
 * fused into the same sweep - the two per-pixel operations align.py applies
 * right after reading: `img_sci -= sky` (align.py:480-481) and
 * `np.equal(wht, 0.0)` (align.py:312, the "DQ" of a drizzled weight image).
 * HBM-bound byte work: every element is read once and written once.
 *
 * raw        the HDU's data block as it is in the file (uploaded by the caller),
 *            16-byte aligned; n elements of |bitpix| / 8 bytes.
 * bitpix     8 (unsigned bytes), 16, 32, 64 (signed), -32, -64 (IEEE).
 * out_kind   SPX_T_F32 / SPX_T_F64: out = T(raw) [* T(bscale)] [+ T(bzero)] [- T(sky)],
 *            each step rounded once in T (no contraction), a step is skipped when
 *            its constant is the identity - numpy's in-place `*=`, `+=`, `-=` on an
 *            array of dtype T, which is how astropy scales (float32 for BITPIX 8 / 16,
 *            float64 for 32 / 64);
 *            SPX_T_I8..I64 of the SAME width as the input (bitpix > 0): integer
 *            payload, bscale must be 1, sky 0, and out = raw + bzero modulo 2^width
 *            (bzero = 2^(width-1) is the FITS convention for unsigned integers:
 *            the stored bits are then the unsigned value).
 * out        16-byte aligned; may alias `raw` when the element widths agree.
 * zero_mask  optional, 16-byte aligned, element type mask_kind (SPX_T_I8 / I32 / I64):
 *            1 where the DECODED value (before the sky subtraction) equals 0.
 * ======================================================================== */
int spx_fits_decode(const void *raw, int64_t n, int bitpix, double bscale, double bzero,
                    double sky, int out_kind, void *out, int mask_kind, void *zero_mask,
                    void *stream);

#ifdef __cplusplus
}
#endif
#endif /* SUBPIXAL_B200_H */
