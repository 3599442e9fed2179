"""Sub-pixel displacement from a 2x-interlaced cross-correlation image.

Same entry point and return types as the reference ``subpixal.cc``
(cc.py:21-95); the work is done by the fused CUDA kernel behind
``spx_cc_fit_stack`` (one launch: normalise, four correlations, interlace,
argmax, 5x5 quadratic fit).
"""
import numpy as np
import torch

from . import batch

__all__ = ["find_displacement"]


def _as_stack(ref_image, phases):
    ref = np.asarray(ref_image)
    ims = [np.asarray(p) for p in phases]
    # cc.py:103-105 - the reference insists on identical shapes
    if not all(tuple(p.shape) == tuple(ref.shape) for p in ims):
        raise ValueError("All cutouts must have same shape.")
    if ref.ndim != 2:
        raise ValueError("Cutouts must be 2-D arrays.")
    out_dtype = np.result_type(ref.dtype, *[p.dtype for p in ims], np.float32)
    if out_dtype not in (np.float32, np.float64):
        out_dtype = np.dtype(np.float64)
    stack = np.empty((1, 5) + ref.shape, dtype=np.float32)
    stack[0, 0] = ref
    for k, p in enumerate(ims):
        stack[0, 1 + k] = p
    return stack, np.dtype(out_dtype)


def find_displacement(ref_image, image00, image10, image01, image11,
                      cc_type='NCC', full_output=False):
    """Displacement of ``image00`` relative to ``ref_image``.

    Parameters and results follow reference cc.py:21-95: ``image10``,
    ``image01`` and ``image11`` are ``image00`` re-sampled on grids displaced by
    half a pixel in x, y and both; ``cc_type`` is ``'CC'``, ``'NCC'`` or
    ``'ZNCC'`` (case-insensitive, anything else acts as ``'CC'``).

    Returns ``(dx, dy)`` or, with ``full_output``, ``(dx, dy, icc, ccs)`` where
    ``icc`` is the ``(2*ny, 2*nx)`` interlaced surface and ``ccs`` the four
    per-phase correlation images ``(cc00, cc10, cc01, cc11)``.

    The kernels compute in float32 (the precision BASELINE.json names); float64
    inputs are narrowed on upload and ``icc`` is returned in the input dtype.
    """
    stack, out_dtype = _as_stack(ref_image, (image00, image10, image01, image11))
    batch.require_cuda()
    dev_stack = torch.from_numpy(stack).cuda()
    shifts, _, _, icc = batch.cc_fit_stack(dev_stack, cc_type, want_icc=full_output)
    dx, dy = (np.float64(v) for v in shifts[0].cpu().numpy())
    if not full_output:
        return dx, dy
    icc = icc[0].cpu().numpy().astype(out_dtype, copy=False)
    # cc.py:121-126 interlaces the reversed per-phase images; undo that
    ccs = tuple(np.ascontiguousarray(icc[b::2, a::2][::-1, ::-1])
                for (a, b) in ((0, 0), (1, 0), (0, 1), (1, 1)))
    return dx, dy, icc, ccs
