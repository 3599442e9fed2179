"""CPU oracle for the subpixal hot path (TEST INFRASTRUCTURE - not product code).

A numpy/scipy restatement of the reference algorithm for
cutout crop -> 4-phase cross-correlation -> interlace -> sub-pixel peak fit.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this module; ``subpixal_b200`` never does.

Parity pin: the reference ships no tests, golden vectors or fixtures
(SURVEY.md section 4), so the pin is "reference source + this image's
numpy 2.3 / scipy 1.18 on identical inputs": ``tests/golden/make_golden.py``
runs the UNMODIFIED reference files (through ``oracle/refload.py``) and
commits its outputs under ``tests/golden/``; ``tests/test_oracle.py`` checks
this restatement against those fixtures (and live against the reference when
``/root/reference`` is mounted).

Third-party arithmetic used exactly like the reference does:
``scipy.signal.fftconvolve`` (cc.py:114-117, scipy un-pinned in setup.py:106,
1.18.1 here) and ``numpy.linalg.lstsq`` (centroid.py:207, numpy>=1.15,
2.3.5 here).
"""
import math

import numpy as np
from scipy.signal import fftconvolve

PHASES = ((0, 0), (1, 0), (0, 1), (1, 1))   # (a, b): x-phase, y-phase


# --------------------------------------------------------------------------
# cc.py
# --------------------------------------------------------------------------
def normalize(ref, phases, zero_mean):
    """cc.py:131-156.  Pooled statistics over the non-zero pixels of the four
    phase images; the reference image is scaled with statistics taken over
    the union of the non-zero masks but applied to every pixel."""
    nz = [p != 0 for p in phases]                               # cc.py:134-137
    pooled = np.concatenate([p[m] for p, m in zip(phases, nz)])  # cc.py:139
    mu = np.mean(pooled) if zero_mean else 0.0                   # cc.py:141
    sigma = np.std(pooled)                                       # cc.py:142
    out = []
    for p, m in zip(phases, nz):                                 # cc.py:144-148
        q = p.copy()
        if zero_mean:
            q[m] -= mu
        q[m] /= sigma
        out.append(q)
    union = np.zeros(ref.shape, dtype=bool)                      # cc.py:150
    for m in nz:
        union |= m
    r = ref.copy()                                               # cc.py:151-154
    if zero_mean:
        r -= np.mean(r[union])
    r /= np.std(r[union])
    return r, out


def build_icc(ref, im00, im10, im01, im11, cc_type="NCC"):
    """cc.py:98-128.  Returns (icc, (cc00, cc10, cc01, cc11))."""
    phases = [im00, im10, im01, im11]
    for p in phases:                                             # cc.py:103-105
        if tuple(p.shape) != tuple(ref.shape):
            raise ValueError("All cutouts must have same shape.")
    kind = cc_type.upper()                                       # cc.py:107
    if kind in ("NCC", "ZNCC"):
        ref, phases = normalize(ref, phases, kind == "ZNCC")     # cc.py:108-111
    ccs = tuple(fftconvolve(ref, p[::-1, ::-1], mode="same")     # cc.py:114-117
                for p in phases)
    ny, nx = ccs[0].shape
    icc = np.empty((2 * ny, 2 * nx), dtype=ccs[0].dtype.type)    # cc.py:121-126
    for (a, b), c in zip(PHASES, ccs):
        icc[b::2, a::2] = c[::-1, ::-1]
    return icc, ccs


def find_displacement(ref_image, image00, image10, image01, image11,
                      cc_type="NCC", full_output=False):
    """cc.py:21-95."""
    icc, ccs = build_icc(ref_image, image00, image10, image01, image11,
                         cc_type)
    xm, ym = find_peak(icc, peak_fit_box=5, peak_search_box="all")  # cc.py:86
    xc = (icc.shape[1] - 1) // 4                                 # cc.py:89-90
    yc = (icc.shape[0] - 1) // 4
    dx = 0.5 * xm - xc                                           # cc.py:92-93
    dy = 0.5 * ym - yc
    return (dx, dy, icc, ccs) if full_output else (dx, dy)


def icc_direct(ref, phases):
    """Independent definition of the interlaced surface by a direct lag sum in
    float64:
        C_ab(ly, lx) = sum_{j,i} ref[j,i] * im_ab[j+ly, i+lx]
        ICC[2k+b, 2l+a] = C_ab(k - yc, l - xc),  yc=(ny-1)//2, xc=(nx-1)//2.
    O(N^4): small cutouts only.  Used to cross-check both the FFT oracle and
    the CUDA path against something that is not an FFT."""
    ref = np.asarray(ref, dtype=np.float64)
    ny, nx = ref.shape
    yc, xc = (ny - 1) // 2, (nx - 1) // 2
    icc = np.zeros((2 * ny, 2 * nx))
    for (a, b), im in zip(PHASES, phases):
        pad = np.zeros((3 * ny, 3 * nx))
        pad[ny:2 * ny, nx:2 * nx] = im
        for k in range(ny):
            for l in range(nx):
                ly, lx = k - yc, l - xc
                win = pad[ny + ly:2 * ny + ly, nx + lx:2 * nx + lx]
                icc[2 * k + b, 2 * l + a] = np.sum(ref * win)
    return icc


# --------------------------------------------------------------------------
# utils.py / centroid.py
# --------------------------------------------------------------------------
def py2round(x):
    """utils.py:144-161 - round half away from zero."""
    if hasattr(x, "__iter__"):
        x = np.asarray(x)
        return np.where(x >= 0.0, np.floor(x + 0.5), np.ceil(x - 0.5))
    return np.floor(x + 0.5) if x >= 0.0 else np.ceil(x - 0.5)


def box_pars(par):
    """centroid.py:239-253."""
    if hasattr(par, "__iter__"):
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
    """centroid.py:18-236, all branches."""
    if (xmax is None) != (ymax is None):                         # :93-96
        raise ValueError("Both 'xmax' and 'ymax' must be either None or not "
                         "None")
    img = np.asarray(image_data, dtype=np.float64)               # :98
    ny, nx = img.shape
    if isinstance(peak_search_box, str):                         # :102-109
        peak_search_box = {"fitbox": peak_fit_box, "off": None,
                           "all": img.shape}.get(peak_search_box,
                                                 peak_search_box)
    expand = False
    if xmax is None:                                             # :111-127
        if mask is None:
            jmax, imax = np.unravel_index(np.argmax(img), img.shape)
        else:
            jj, ii = np.nonzero(mask)          # row-major like np.indices[mask]
            k = np.argmax(img[mask])
            imax, jmax = ii[k], jj[k]
        coord = (float(imax), float(jmax))
    else:                                                        # :129-156
        imax, jmax = int(py2round(xmax)), int(py2round(ymax))
        coord = (xmax, ymax)
        if peak_search_box is not None:
            sbx, sby = box_pars(peak_search_box)
            x1 = max(0, imax - sbx // 2)
            x2 = min(nx, x1 + sbx)
            y1 = max(0, jmax - sby // 2)
            y2 = min(ny, y1 + sby)
            if x1 < x2 and y1 < y2:
                sub = img[y1:y2, x1:x2]
                jmax, imax = np.unravel_index(np.argmax(sub), sub.shape)
                imax += x1
                jmax += y1
                coord = (float(imax), float(jmax))
            expand = (sbx != nx or sby != ny)
    wx, wy = box_pars(peak_fit_box)                              # :158
    if wx * wy < 6:                                              # :160-162
        return coord
    x1 = max(0, imax - wx // 2)                                  # :165-168
    x2 = min(nx, x1 + wx)
    y1 = max(0, jmax - wy // 2)
    y2 = min(ny, y1 + wy)
    if imax == x1 or imax == x2 or jmax == y1 or jmax == y2:     # :171-172
        return (float(imax), float(jmax))
    if x2 - x1 < wx:                                             # :175-184
        if x1 == 0:
            x2 = min(nx, x1 + wx)
        if x2 == nx:
            x1 = max(0, x2 - wx)
    if y2 - y1 < wy:
        if y1 == 0:
            y2 = min(ny, y1 + wy)
        if y2 == ny:
            y1 = max(0, y2 - wy)
    if (x2 - x1) * (y2 - y1) < 6:                                # :186-188
        return coord
    xs, ys = np.meshgrid(np.arange(x1, x2), np.arange(y1, y2))   # :191-197
    xs, ys = xs.ravel(), ys.ravel()
    design = np.stack([np.ones_like(xs), xs, ys, xs * ys, xs * xs, ys * ys],
                      axis=1)
    vals = img[y1:y2, x1:x2].ravel()
    if mask is not None:                                         # :198-204
        good = np.asarray(mask)[y1:y2, x1:x2].ravel()
        design, vals = design[good], vals[good]
        if vals.size < 6:
            return coord

    def whole_image():
        return find_peak(img, peak_fit_box=(wx, wy), mask=mask)

    coef = np.linalg.lstsq(design, vals, rcond=None)[0]          # :207
    _, c10, c01, c11, c20, c02 = coef
    det = 4 * c02 * c20 - c11 ** 2                               # :216-225
    if det <= 0 or ((c20 > 0.0 and c02 >= 0.0) or (c20 >= 0.0 and c02 > 0.0)):
        return whole_image() if expand else ((x1 + x2) / 2.0, (y1 + y2) / 2.0)
    xm = (c01 * c11 - 2.0 * c02 * c10) / det                     # :227-228
    ym = (c10 * c11 - 2.0 * c01 * c20) / det
    if 0.0 < xm < nx - 1.0 and 0.0 < ym < ny - 1.0:              # :230-236
        return (xm, ym)
    return whole_image() if expand else coord


# --------------------------------------------------------------------------
# cutout.py (crop / mask arithmetic only; WCS methods stay in astropy)
# --------------------------------------------------------------------------
class NoOverlap(ValueError):
    pass


class PartialOverlap(ValueError):
    pass


def bitfield_bad_mask(dq, good_flags):
    """stsci.tools.bitmask.bitfield_to_boolean_mask(dq, ignore_flags=good_flags,
    good_mask_value=False) for integer flags, as used at cutout.py:191, 287,
    422 and documented at cutout.py:77-109.  True = bad pixel."""
    dq = np.asarray(dq)
    if good_flags is None:
        return np.zeros(dq.shape, dtype=bool)
    width = 8 * dq.dtype.itemsize
    inv = (~int(good_flags)) & ((1 << width) - 1)           # two's complement, dtype width
    bits = dq.view(np.dtype("u%d" % dq.dtype.itemsize))
    return (bits & np.array(inv, dtype=bits.dtype)) != 0


def crop(data, blc, trc, dq=None, weight=None, mode="strict",
         fillval=np.nan):
    """Cutout.__init__ arithmetic, cutout.py:689-797.  Returns a dict with
    data, mask (True = bad), dq, weight, extraction/insertion slices."""
    ny, nx = data.shape
    if mode not in ("strict", "fill"):
        raise ValueError("Argument 'mode' must be either 'strict' or 'fill'.")
    if trc is None:
        trc = (nx - 1, ny - 1)
    if blc[0] >= nx or blc[1] >= ny or trc[0] < 0 or trc[1] < 0:   # :713-716
        raise NoOverlap("Cutout's extraction box does not overlap image "
                        "data array.")
    if trc[0] < blc[0] or trc[1] < blc[1]:                         # :719-722
        raise ValueError("Ill-formed extraction box")
    if mode == "strict" and (blc[0] < 0 or blc[1] < 0 or
                             trc[0] >= nx or trc[1] >= ny):        # :724-729
        raise PartialOverlap("Cutout's extraction box only partially "
                             "overlaps image data array.")
    h, w = trc[1] - blc[1] + 1, trc[0] - blc[0] + 1               # :1024-1031
    out = np.full((h, w), fillval, dtype=data.dtype)               # :737-739
    bad = np.ones((h, w), dtype=bool)
    ex = np.s_[max(0, blc[1]):min(ny - 1, trc[1]) + 1,
               max(0, blc[0]):min(nx - 1, trc[0]) + 1]             # :742-747
    ins = np.s_[ex[0].start - blc[1]:ex[0].stop - blc[1],
                ex[1].start - blc[0]:ex[1].stop - blc[0]]
    out[ins] = data[ex]                                            # :750-752
    bad[ins] = False
    bad |= ~np.isfinite(out)                                       # :755
    res = {"data": out, "mask": bad, "extraction_slice": ex,
           "insertion_slice": ins, "dq": None, "weight": None}
    for key, arr in (("dq", dq), ("weight", weight)):              # :758-779
        if arr is not None:
            if arr.shape != (ny, nx):
                raise ValueError("shape mismatch for " + key)
            z = np.zeros((h, w), dtype=arr.dtype)
            z[ins] = arr[ex]
            res[key] = z
    return res


# --------------------------------------------------------------------------
# Deterministic synthetic inputs
# --------------------------------------------------------------------------
def gaussian_star(n, x0, y0, sigma=2.0, amp=1000.0, nx=None):
    ny, nx = n, (n if nx is None else nx)
    y, x = np.mgrid[:ny, :nx].astype(np.float64)
    return (amp * np.exp(-((x - x0) ** 2 + (y - y0) ** 2)
                         / (2.0 * sigma * sigma))).astype(np.float32)


def known_shift_case(n, shift=(0.37, -0.21), sigma=2.0):
    """Noise-free config-1 case: ref at the centre, image_ab sampled on grids
    displaced by +a/2, +b/2 px (features move left/down)."""
    c = (n - 1) / 2.0
    ref = gaussian_star(n, c, c, sigma)
    ims = [gaussian_star(n, c + shift[0] - a / 2.0, c + shift[1] - b / 2.0,
                         sigma) for a, b in PHASES]
    return ref, ims


def random_pairs(n_pairs, ny, nx=None, seed=0, noise=1.0, masked=False):
    """Noisy random-source pairs, stack layout [n, 5, ny, nx] float32:
    slot 0 = ref_image, slots 1..4 = image00, image10, image01, image11."""
    nx = ny if nx is None else nx
    rng = np.random.default_rng(seed)
    out = np.empty((n_pairs, 5, ny, nx), dtype=np.float32)
    truth = np.empty((n_pairs, 2))
    y, x = np.mgrid[:ny, :nx].astype(np.float64)
    for p in range(n_pairs):
        s = rng.uniform(1.5, min(4.0, max(1.6, min(ny, nx) / 8.0)))
        q = rng.uniform(0.5, 1.0)
        amp = math.exp(rng.uniform(math.log(50.0), math.log(5000.0)))
        cx = (nx - 1) / 2.0 + rng.uniform(-1.5, 1.5)
        cy = (ny - 1) / 2.0 + rng.uniform(-1.5, 1.5)
        sx, sy = rng.uniform(-0.9, 0.9, size=2)
        truth[p] = (sx, sy)

        def render(x0, y0):
            return amp * np.exp(-((x - x0) ** 2 / (2 * s * s)
                                  + (y - y0) ** 2 / (2 * (s * q) ** 2)))
        out[p, 0] = render(cx, cy) + noise * rng.standard_normal((ny, nx))
        for k, (a, b) in enumerate(PHASES):
            im = render(cx + sx - a / 2.0, cy + sy - b / 2.0)
            if masked:
                rad = 0.4 * min(ny, nx)
                im = np.where((x - cx) ** 2 + (y - cy) ** 2 <= rad * rad,
                              im, 0.0)
            out[p, 1 + k] = im
    return out, truth
