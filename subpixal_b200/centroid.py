"""Peak finding by a quadratic fit around the brightest pixel.

Same signature as the reference ``subpixal.centroid.find_peak``
(centroid.py:18-236).  Argument handling stays on the host; the argmax, the
search-box scan and the least-squares fit run in the ``spx_find_peak`` kernel.
"""
import numpy as np
import torch

from . import _lib as L
from . import batch
from .utils import native_array

__all__ = ["find_peak"]

_RETRY = L.S_NO_MAXIMUM | L.S_FIT_OUTSIDE | L.S_FIT_NONFINITE | L.S_SINGULAR


def _box(par):
    """centroid.py:239-253: scalar or pair -> (wx, wy)."""
    if hasattr(par, '__iter__'):
        if len(par) != 2:
            raise TypeError("Box specification must be either a scalar or "
                            "an iterable with two elements.")
        wx, wy = int(par[0]), int(par[1])
    else:
        wx = wy = int(par)
    if wx < 1 or wy < 1:
        raise ValueError("Box dimensions must be positive integer numbers.")
    return wx, wy


def find_peak(image_data, xmax=None, ymax=None, peak_fit_box=5,
              peak_search_box=None, mask=None):
    """Location ``(x, y)`` of the peak of ``image_data``.

    A 2-D quadratic is fitted to the ``peak_fit_box`` around the brightest pixel
    (or around ``(xmax, ymax)``, optionally refined by a brute-force search in
    ``peak_search_box``: ``'all'``, ``'off'``, ``'fitbox'``, an int or a
    ``(width, height)`` pair); ``mask`` marks good pixels with ``True``.
    Fall-backs (integer position at the box edge, box centre when the fit has
    no maximum, original coordinate when the fit leaves the image) are those of
    reference centroid.py:160-236.
    """
    if (xmax is None) != (ymax is None):                     # centroid.py:93-96
        raise ValueError("Both 'xmax' and 'ymax' must be either None or not "
                         "None")
    img = native_array(image_data)                           # FITS arrays are big-endian
    if img.dtype != np.float32:                              # centroid.py:98
        img = np.asarray(img, dtype=np.float64)
    if img.ndim != 2:
        raise ValueError("image_data must be a 2-D array")
    ny, nx = img.shape
    if isinstance(peak_search_box, str):                     # centroid.py:102-109
        peak_search_box = {'fitbox': peak_fit_box, 'off': None,
                           'all': img.shape}.get(peak_search_box, peak_search_box)
    wx, wy = _box(peak_fit_box)
    if wx > L.MAX_FIT_BOX or wy > L.MAX_FIT_BOX:
        raise ValueError("peak_fit_box larger than %d is not supported" % L.MAX_FIT_BOX)
    search, expand = (0, 0), False
    if xmax is not None and peak_search_box is not None:
        search = _box(peak_search_box)
        expand = (search[0] != nx or search[1] != ny)        # centroid.py:153
    good = None
    if mask is not None:
        good = np.asarray(mask, dtype=bool)
        if good.shape != img.shape:
            raise ValueError("mask shape must match image_data")
        if xmax is None and not good.any():
            raise ValueError("attempt to get argmax of an empty sequence")

    batch.require_cuda()
    surf = torch.from_numpy(np.ascontiguousarray(img)).cuda()[None]
    dmask = torch.from_numpy(good.view(np.uint8)).cuda()[None] if good is not None else None
    guess = (torch.tensor([[float(xmax), float(ymax)]], dtype=torch.float64)
             if xmax is not None else None)
    xy, _, status = batch.find_peak_batch(surf, dmask, guess, search, (wx, wy))
    st = int(status[0])
    if expand and (st & _RETRY):
        # centroid.py:210-212, 221-223, 232-234: retry on the whole image
        xy, _, status = batch.find_peak_batch(surf, dmask, None, (0, 0), (wx, wy))
    x, y = xy[0].cpu().numpy()
    if xmax is not None and peak_search_box is None and (int(status[0]) & ~L.S_PEAK_EDGE) != 0 \
            and (x, y) == (float(xmax), float(ymax)):
        return (xmax, ymax)          # hand the caller's own objects back
    return (float(x), float(y))
