"""GPU blot of image cutouts - drop-in for ``subpixal.blot`` (reference blot.py).

``blot_cutout(source_cutout, target_cutout, interp='poly5', sinscl=1.0,
wcsmap=None)`` keeps the reference's signature, return type (a deep copy of the
target cutout whose ``data`` holds the blotted source) and errors
(blot.py:79-155); ``BlotWCSMap`` is the reference's default coordinate map
(blot.py:21-76).  ``blot_phases`` is the batched form used by
``align.find_linear_fit``: every source at the four half-pixel phases of
align.py:661-679 in ONE launch of ``spx_blot_phases``.

Differences from the reference, by construction:

* the reference hands ``wcsmap`` to drizzlepac's C code, which calls it back for
  every output pixel; here ``wcsmap`` is sampled on a 3x3 grid of the target
  cutout (corners, edge mid-points, centre) and the interpolating biquadratic is
  evaluated on the device (float64).  Over a cutout of <= 256 px the cubic terms
  of HST distortion models are < 1e-4 px.
* only ``interp`` in {'nearest', 'linear', 'poly3', 'poly5'} is implemented on the GPU
  ('poly5' is the reference's default and the only one align.py uses); the other
  drizzle kernels raise ``NotImplementedError``.
* drizzlepac is an un-vendored dependency that is absent here: the interpolator
  restates its published algorithm, parity with ``cdriz.tblot`` is UNPINNED
  (see oracle/blot_oracle.py and DESIGN.md).
"""
import copy

import numpy as np
import torch

from . import batch

__all__ = ['BlotWCSMap', 'blot_cutout', 'blot_phases', 'fit_pixel_map', 'fit_pixel_maps',
           'pixel_ratio']

_SUPPORTED = ('nearest', 'linear', 'poly3', 'poly5')
_DRIZZLE_ONLY = ('spline3', 'sinc', 'lanczos3', 'lanczos5')


class BlotWCSMap(object):
    """Maps pixel coordinates of ``source_cutout`` to pixel coordinates of
    ``target_cutout`` through the sky; 1-based coordinates in and out
    (reference blot.py:21-76)."""

    def __init__(self, source_cutout, target_cutout):
        self._source_cutout = source_cutout
        self._target_cutout = target_cutout

    def __call__(self, x, y):
        sky = self._source_cutout.pix2world(x, y, origin=1)
        return self._target_cutout.world2pix(*sky, origin=1)


def fit_pixel_map(wcsmap, nx, ny):
    """``[20]`` float64 coefficients (xc, yc, a[3][3], b[3][3]) of the biquadratic
    that interpolates ``wcsmap`` (1-based x, y -> 1-based xs, ys) on the 3x3 grid
    of an ``nx`` x ``ny`` target; expressed in 0-based pixels on both sides."""
    xc, yc = 0.5 * (nx - 1), 0.5 * (ny - 1)
    hx, hy = max(xc, 0.5), max(yc, 0.5)
    uu, vv = np.meshgrid(np.array([-hx, 0.0, hx]), np.array([-hy, 0.0, hy]), indexing="ij")
    xs, ys = wcsmap((uu + xc).ravel() + 1.0, (vv + yc).ravel() + 1.0)
    xs = np.asarray(xs, np.float64).reshape(3, 3) - 1.0
    ys = np.asarray(ys, np.float64).reshape(3, 3) - 1.0

    def diff(f, h, axis):     # samples at (-h, 0, h) -> coefficients of 1, t, t^2
        fm, f0, fp = np.moveaxis(f, axis, 0)
        c = np.stack([f0, (fp - fm) / (2.0 * h), ((fp - f0) - (f0 - fm)) / (2.0 * h * h)])
        return np.moveaxis(c, 0, axis)

    a = diff(diff(xs, hx, 0), hy, 1)
    b = diff(diff(ys, hx, 0), hy, 1)
    return np.concatenate([[xc, yc], a.ravel(), b.ravel()])


def fit_pixel_maps(frame_map, target_boxes, source_boxes):
    """Vectorised ``fit_pixel_map`` for n cutouts of two full frames: ``frame_map(x, y)``
    maps 0-based pixels of the target (image) frame to 0-based pixels of the source
    (drizzled) frame and accepts arrays; boxes are ``[n, 4]`` (x1, y1, x2, y2), trc
    inclusive.  One call of ``frame_map`` with 9n points.  Returns ``[n, 20]``."""
    tb = np.asarray(target_boxes, dtype=np.float64).reshape(-1, 4)
    sb = np.asarray(source_boxes, dtype=np.float64).reshape(-1, 4)
    n = tb.shape[0]
    xc = 0.5 * (tb[:, 2] - tb[:, 0])
    yc = 0.5 * (tb[:, 3] - tb[:, 1])
    hx, hy = np.maximum(xc, 0.5), np.maximum(yc, 0.5)
    # node (p, q) = (xc + (p-1) hx, yc + (q-1) hy); all arrays are [9, n] with n contiguous
    px = np.stack([xc + tb[:, 0] + (p - 1) * hx for p in range(3) for q in range(3)])
    py = np.stack([yc + tb[:, 1] + (q - 1) * hy for p in range(3) for q in range(3)])
    gx, gy = frame_map(px.ravel(), py.ravel())
    out = np.empty((n, 20), np.float64)
    out[:, 0], out[:, 1] = xc, yc
    for base, g, origin in ((2, gx, sb[:, 0]), (11, gy, sb[:, 1])):
        f = np.asarray(g, np.float64).reshape(3, 3, n) - origin           # [p, q, n]
        # coefficients of 1, u, u^2 along p (exact for linear maps), then of 1, v, v^2 along q
        cu = (f[1], (f[2] - f[0]) / (2.0 * hx), ((f[2] - f[1]) - (f[1] - f[0])) / (2.0 * hx * hx))
        for p in range(3):
            c = cu[p]                                                      # [q, n]
            out[:, base + 3 * p + 0] = c[1]
            out[:, base + 3 * p + 1] = (c[2] - c[0]) / (2.0 * hy)
            out[:, base + 3 * p + 2] = ((c[2] - c[1]) - (c[1] - c[0])) / (2.0 * hy * hy)
    return out


_stwcs_distortion = False          # False: not looked for yet; None: absent


def _orthogonal_cd():
    """``stwcs.distortion.utils.make_orthogonal_cd`` or None (looked up once: a failing
    import costs about a millisecond EVERY time it is retried)."""
    global _stwcs_distortion
    if _stwcs_distortion is False:
        try:
            from stwcs import distortion
            _stwcs_distortion = distortion.utils.make_orthogonal_cd
        except ImportError:
            _stwcs_distortion = None
    return _stwcs_distortion


def pixel_ratio(source_cutout, target_cutout, _cache=None):
    """blot.py:123-133: source pixel scale over the pixel scale of the target's
    orthogonalised linear WCS (``stwcs.distortion.utils.make_orthogonal_cd``) when
    stwcs is importable, else the ratio of the cutouts' own pixel scales.  ``_cache``:
    a dict shared by the pairs of one batch (cutouts of one image share their WCS)."""
    key = (id(source_cutout.wcs), id(target_cutout.wcs))
    if _cache is not None and key in _cache:
        return _cache[key]
    ortho = _orthogonal_cd()
    if ortho is not None:
        pr = float(source_cutout.wcs.pscale / ortho(target_cutout.wcs).pscale)
    else:
        pr = float(source_cutout.pscale / target_cutout.pscale)
    if _cache is not None:
        _cache[key] = pr
    return pr


def _cutout_pixel_maps(source_cutouts, target_cutouts):
    """``fit_pixel_map(BlotWCSMap(target, source), nx, ny)`` for every pair, with ONE call
    of each WCS per group of pairs that share their two WCS objects (the cutouts of one
    image do): the same element-wise arithmetic in the same order as the per-pair path
    (Cutout.pix2world / world2pix: cutout.py:1070-1145), so the maps are identical
    whenever the WCS functions are element-wise."""
    n = len(source_cutouts)
    out = np.empty((n, 20), np.float64)
    groups = {}
    for k, (sct, tct) in enumerate(zip(source_cutouts, target_cutouts)):
        groups.setdefault((id(tct.wcs), id(sct.wcs)), []).append(k)
    for idx in groups.values():
        tcts = [target_cutouts[k] for k in idx]
        scts = [source_cutouts[k] for k in idx]
        m = len(idx)
        nxy = np.array([[int(v) for v in t.naxis] for t in tcts], dtype=np.float64)     # (nx, ny)
        xc, yc = 0.5 * (nxy[:, 0] - 1), 0.5 * (nxy[:, 1] - 1)
        hx, hy = np.maximum(xc, 0.5), np.maximum(yc, 0.5)
        # node (p, q) of pair i: u = (p-1) hx, v = (q-1) hy ("ij" meshgrid order of fit_pixel_map)
        sgn = np.array([-1.0, 0.0, 1.0])
        uu = (sgn[None, :, None] * hx[:, None, None]) * np.ones((1, 1, 3))            # [m, p, q]
        vv = (sgn[None, None, :] * hy[:, None, None]) * np.ones((1, 3, 1))
        x1 = (uu + xc[:, None, None]).reshape(m, 9) + 1.0                               # 1-based cutout px
        y1 = (vv + yc[:, None, None]).reshape(m, 9) + 1.0
        toff = np.array([[t.blc[0] - t.dx, t.blc[1] - t.dy] for t in tcts], dtype=np.float64)
        soff = np.array([[c.blc[0] - c.dx, c.blc[1] - c.dy] for c in scts], dtype=np.float64)
        ra, dec = tcts[0].wcs.all_pix2world((x1 + toff[:, 0:1]).ravel(), (y1 + toff[:, 1:2]).ravel(), 1)
        sc0 = scts[0]
        px, py = sc0.wcs.all_world2pix(ra, dec, 1, accuracy=sc0.DEFAULT_ACCURACY,
                                       maxiter=sc0.DEFAULT_MAXITER, quiet=sc0.DEFAULT_QUIET)
        xs = (np.asarray(px, np.float64).reshape(m, 9) - soff[:, 0:1]).reshape(m, 3, 3) - 1.0
        ys = (np.asarray(py, np.float64).reshape(m, 9) - soff[:, 1:2]).reshape(m, 3, 3) - 1.0

        def diff(f, h, axis):     # samples at (-h, 0, h) along `axis` -> coefficients of 1, t, t^2
            fm, f0, fp = np.moveaxis(f, axis, 0)
            hh = h.reshape((m,) + (1,) * (fm.ndim - 1))
            c = np.stack([f0, (fp - fm) / (2.0 * hh), ((fp - f0) - (f0 - fm)) / (2.0 * hh * hh)])
            return np.moveaxis(c, 0, axis)

        a = diff(diff(xs, hx, 1), hy, 2)
        b = diff(diff(ys, hx, 1), hy, 2)
        out[idx, 0], out[idx, 1] = xc, yc
        out[idx, 2:11] = a.reshape(m, 9)
        out[idx, 11:20] = b.reshape(m, 9)
    return out


def _check(source_cutout, target_cutout, interp):
    if len(source_cutout.naxis) != 2 or len(target_cutout.naxis) != 2:
        raise ValueError("Cutouts must be 2D images.")                # blot.py:109-110
    if interp in _DRIZZLE_ONLY:
        raise NotImplementedError(
            "interp=%r is a drizzlepac kernel that has no GPU implementation here; "
            "use one of %s" % (interp, ", ".join(_SUPPORTED)))
    if interp not in _SUPPORTED:
        raise ValueError("Unknown interpolation %r" % (interp,))


def _flux_scale(source_cutout, target_cutout, pix_ratio):
    """blot.py:134-150 folded into one factor: counts -> rate on the way in,
    tblot's ``ef / scale**2`` with ef = target exptime, rate on the way out."""
    sc = 1.0
    if source_cutout.data_units == 'counts':
        sc /= source_cutout.exptime
    sc *= target_cutout.exptime / (pix_ratio * pix_ratio)
    if target_cutout.data_units == 'rate':
        sc /= target_cutout.exptime
    return sc


def blot_phases(source_cutouts, target_cutouts, phases=batch.PHASES_00_10_01_11,
                interp='poly5', wcsmaps=None, pix_ratios=None, device=None, as_stack=False):
    """Blot every source cutout onto its target cutout's grid at all ``phases``
    in one launch.  Returns a list of float32 CUDA tensors ``[nph, ny, nx]`` (views
    into one buffer), one per pair.  ``wcsmaps`` / ``pix_ratios``: optional
    per-pair overrides (callables like ``BlotWCSMap`` / floats).

    ``as_stack=True`` (four phases): returns ``(stack, offsets, shapes)`` laid out
    for ``batch.cc_fit_ragged`` - pair k is ``[5, ny, nx]`` at ``offsets[k]`` with the
    blots in slots 1..4 and slot 0 (``ref_image``) left for the caller to fill."""
    batch.require_cuda()
    source_cutouts, target_cutouts = list(source_cutouts), list(target_cutouts)
    n = len(source_cutouts)
    if len(target_cutouts) != n:
        raise ValueError("The number of source cutouts must match the number of target cutouts.")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    maps = np.empty((n, 20), np.float64)
    scale = np.empty(n, np.float32)
    s_shp = np.empty((n, 2), np.int64)
    t_shp = np.empty((n, 2), np.int64)
    chunks = []
    if wcsmaps is None and n:
        for sct, tct in zip(source_cutouts, target_cutouts):
            _check(sct, tct, interp)
        maps[:] = _cutout_pixel_maps(source_cutouts, target_cutouts)  # blot.py:119-120, batched
    ratio_cache = {}
    for k, (sct, tct) in enumerate(zip(source_cutouts, target_cutouts)):
        _check(sct, tct, interp)
        nx, ny = (int(v) for v in tct.naxis)
        if wcsmaps is not None:
            wm = wcsmaps[k]
            if wm is None:
                wm = BlotWCSMap(tct, sct)                             # blot.py:119-120
            maps[k] = fit_pixel_map(wm, nx, ny)
        pr = pix_ratios[k] if pix_ratios is not None else pixel_ratio(sct, tct, ratio_cache)
        scale[k] = _flux_scale(sct, tct, float(pr))
        data = np.ascontiguousarray(sct.data, dtype=np.float32)       # blot.py:135
        s_shp[k] = data.shape
        t_shp[k] = (ny, nx)
        chunks.append(data.ravel())
    sizes = s_shp[:, 0] * s_shp[:, 1]
    s_off = np.concatenate([[0], np.cumsum(sizes)]) if n else np.zeros(1, np.int64)
    flat = np.concatenate(chunks) if n else np.zeros(0, np.float32)
    nph = len(tuple(phases))
    if as_stack:
        if nph != 4:
            raise ValueError("as_stack needs the four phases of find_displacement")
        tiles = t_shp[:, 0] * t_shp[:, 1]
        p_off = np.concatenate([[0], np.cumsum(5 * tiles)]).astype(np.int64)
        with torch.cuda.device(dev):
            stack = torch.empty(int(p_off[-1]), dtype=torch.float32, device=dev)
            src = torch.from_numpy(flat).to(dev)
            batch.blot_phases(src, s_off[:-1], s_shp, maps, t_shp, phases=phases, scale=scale,
                              interp=interp, misval=0.0, out=stack, out_offsets=p_off[:-1] + tiles)
        return stack, p_off[:-1], t_shp
    with torch.cuda.device(dev):
        src = torch.from_numpy(flat).to(dev)
        out, o_off = batch.blot_phases(src, s_off[:-1], s_shp, maps, t_shp, phases=phases,
                                       scale=scale, interp=interp, misval=0.0)
    res = []
    for k in range(n):
        ny, nx = int(t_shp[k, 0]), int(t_shp[k, 1])
        res.append(out[int(o_off[k]):int(o_off[k + 1])].view(nph, ny, nx))
    return res


def blot_cutout(source_cutout, target_cutout, interp='poly5', sinscl=1.0,
                wcsmap=None):
    """Blot ``source_cutout`` onto the grid of ``target_cutout`` (reference
    blot.py:79-155).  Returns a deep copy of ``target_cutout`` whose ``data`` is
    the blotted image; ``sinscl`` only matters for the 'sinc' kernel, which is
    not available on the GPU."""
    res = blot_phases([source_cutout], [target_cutout], phases=((0.0, 0.0),), interp=interp,
                      wcsmaps=[wcsmap])
    out_cutout = copy.deepcopy(target_cutout)                          # blot.py:152-155
    out_cutout.data[:, :] = res[0][0].cpu().numpy()
    return out_cutout
