"""CPU oracle for the blot row - TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU arms may import
this module; the product (``subpixal_b200``) never does.

**Parity unpinned.**  The reference's blot (reference blot.py:79-155, called four
times per source at align.py:661-679) is a thin wrapper around
``drizzlepac.cdriz.tblot`` - C code of the un-vendored third-party dependency
``drizzlepac`` (setup.py:107, no version pin; its ``src/cdrizzleblot.c``:
``doblot`` / ``interpolate_poly5`` / ``ii_bipoly5``).  drizzlepac is absent from
/root/reference and from this image and there is no network, so neither the C
source nor its outputs were available: what follows restates the *published*
algorithm (Everett central-difference form of the 6x6 quintic interpolator that
drizzle inherited from IRAF's ``ii_bipoly5``, edge rows/columns extended by the
linear reflection ``2*f(edge) - f(mirror)``, pixels that map outside the source
set to ``misval``, flux scaled by ``ef / pix_ratio**2``) from the call site's
arguments (blot.py:131-147: ``interp='poly5'``, ``misval=0``, ``kscale=1``,
``xmin=ymin=1``, ``xmax,ymax=source naxis``, 1-based ``wcsmap``).  No golden
vector of the reference exists for this path; the GPU kernel is checked
bit-for-bit against THIS restatement and through size-independent properties
(polynomial reproduction, identity map, shift consistency).

All arithmetic is float32 with one rounding per operation (no fused
multiply-add), coordinates are evaluated in float64 and rounded to float32
once - the CUDA kernel uses the same operation order with explicit
round-to-nearest intrinsics, so the two agree bit-for-bit.
"""
import numpy as np

F = np.float32


# --------------------------------------------------------------------------
# coordinate map: tensor-product biquadratic in centred target pixel coordinates
#   coef = [xc, yc, a00..a22 (9), b00..b22 (9)]   (20 float64)
#   xs = sum_pq a[p][q] * u**p * v**q,  u = x - xc, v = y - yc   (0-based pixels)
# --------------------------------------------------------------------------
def fit_map(fn, nx, ny):
    """Interpolating biquadratic of ``fn(x, y) -> (xs, ys)`` (0-based pixel
    coordinates in, 0-based source pixel coordinates out) through the 3x3 grid of
    the target cutout's corners, edge mid-points and centre."""
    xc, yc = 0.5 * (nx - 1), 0.5 * (ny - 1)
    hx, hy = max(xc, 0.5), max(yc, 0.5)
    us = np.array([-hx, 0.0, hx])
    vs = np.array([-hy, 0.0, hy])
    uu, vv = np.meshgrid(us, vs, indexing="ij")              # [p-node, q-node]
    xs, ys = fn((uu + xc).ravel(), (vv + yc).ravel())
    xs = np.asarray(xs, np.float64).reshape(3, 3)
    ys = np.asarray(ys, np.float64).reshape(3, 3)

    def diff(f, h, axis):   # samples at (-h, 0, h) -> coefficients of 1, t, t^2 (exact for linear data)
        fm, f0, fp = np.moveaxis(f, axis, 0)
        c = np.stack([f0, (fp - fm) / (2.0 * h), ((fp - f0) - (f0 - fm)) / (2.0 * h * h)])
        return np.moveaxis(c, 0, axis)
    a = diff(diff(xs, hx, 0), hy, 1)                          # a[p][q]
    b = diff(diff(ys, hx, 0), hy, 1)
    return np.concatenate([[xc, yc], a.ravel(), b.ravel()])


def eval_map(coef, x, y):
    """(xs, ys) in float64, Horner in v then in u, one rounding per operation."""
    coef = np.asarray(coef, np.float64)
    u = np.asarray(x, np.float64) - coef[0]
    v = np.asarray(y, np.float64) - coef[1]
    out = []
    for base in (2, 11):
        c = coef[base:base + 9].reshape(3, 3)
        r = [(c[p, 2] * v + c[p, 1]) * v + c[p, 0] for p in range(3)]
        out.append((r[2] * u + r[1]) * u + r[0])
    return out[0], out[1]


# --------------------------------------------------------------------------
# interpolators (float32, drizzle's operation order)
# --------------------------------------------------------------------------
def _everett(fm2, fm1, f0, f1, f2, f3, s):
    """Quintic through six equally spaced samples (positions -2..3) at 0 <= s < 1."""
    one, two, four, six = F(1), F(2), F(4), F(6)
    s2 = s * s
    t = one - s
    t2 = t * t
    cd20 = F(1.0 / 6.0) * ((f1 - two * f0) + fm1)
    cd21 = F(1.0 / 6.0) * ((f2 - two * f1) + f0)
    cd40 = F(1.0 / 120.0) * ((((fm2 - four * fm1) + six * f0) - four * f1) + f2)
    cd41 = F(1.0 / 120.0) * ((((fm1 - four * f0) + six * f1) - four * f2) + f3)
    return (s * (f1 + (s2 - one) * (cd21 + (s2 - four) * cd41)) +
            t * (f0 + (t2 - one) * (cd20 + (t2 - four) * cd40)))


def _extended(data, jj, ii):
    """Source samples with the linear edge extension 2*f(edge) - f(mirror),
    columns first and rows on the extended columns (float32)."""
    h, w = data.shape

    def fold(k, n):
        lo, hi = k < 0, k >= n
        km = np.where(lo, -k, np.where(hi, 2 * (n - 1) - k, k))
        ke = np.where(lo, 0, np.where(hi, n - 1, k))
        # a 6-wide stencil on a tiny image can fold past the far edge: clamp
        return np.clip(km, 0, n - 1), ke, lo | hi

    im, ie, xo = fold(ii, w)
    jm, je, yo = fold(jj, h)

    def col(j):             # value at (row j [in range], extended column)
        direct = data[j, im]
        return np.where(xo, F(2) * data[j, ie] - direct, direct)
    direct = col(jm)
    return np.where(yo, F(2) * col(je) - direct, direct).astype(F)


def interp_poly5(data, x, y):
    """6x6 quintic interpolation of float32 ``data`` at 0-based float32 (x, y)."""
    data = np.asarray(data, F)
    x = np.asarray(x, F)
    y = np.asarray(y, F)
    ix = x.astype(np.int64)                  # C cast: truncation (x >= 0 here)
    iy = y.astype(np.int64)
    sx = x - ix.astype(F)
    sy = y - iy.astype(F)
    z = []
    for dj in range(-2, 4):
        f = [_extended(data, iy + dj, ix + di) for di in range(-2, 4)]
        z.append(_everett(*f, sx))
    return _everett(*z, sy)


def _everett3(fm1, f0, f1, f2, s):
    """Cubic through four equally spaced samples (positions -1..2) at 0 <= s < 1
    (IRAF ii_bipoly3: Everett's formula truncated after the second differences)."""
    one, two = F(1), F(2)
    s2 = s * s
    t = one - s
    t2 = t * t
    cd20 = F(1.0 / 6.0) * ((f1 - two * f0) + fm1)
    cd21 = F(1.0 / 6.0) * ((f2 - two * f1) + f0)
    return s * (f1 + (s2 - one) * cd21) + t * (f0 + (t2 - one) * cd20)


def interp_poly3(data, x, y):
    """4x4 cubic interpolation of float32 ``data`` at 0-based float32 (x, y)."""
    data = np.asarray(data, F)
    x = np.asarray(x, F)
    y = np.asarray(y, F)
    ix = x.astype(np.int64)
    iy = y.astype(np.int64)
    sx = x - ix.astype(F)
    sy = y - iy.astype(F)
    z = []
    for dj in range(-1, 3):
        f = [_extended(data, iy + dj, ix + di) for di in range(-1, 3)]
        z.append(_everett3(*f, sx))
    return _everett3(*z, sy)


def interp_linear(data, x, y):
    data = np.asarray(data, F)
    h, w = data.shape
    x = np.asarray(x, F)
    y = np.asarray(y, F)
    ix = x.astype(np.int64)
    iy = y.astype(np.int64)
    sx = x - ix.astype(F)
    sy = y - iy.astype(F)
    tx = F(1) - sx
    ty = F(1) - sy
    i1 = np.minimum(ix + 1, w - 1)
    j1 = np.minimum(iy + 1, h - 1)
    return ((tx * ty) * data[iy, ix] + (sx * ty) * data[iy, i1] +
            (tx * sy) * data[j1, ix] + (sx * sy) * data[j1, i1]).astype(F)


def interp_nearest(data, x, y):
    data = np.asarray(data, F)
    ix = (np.asarray(x, F) + F(0.5)).astype(np.int64)
    iy = (np.asarray(y, F) + F(0.5)).astype(np.int64)
    h, w = data.shape
    return data[np.minimum(iy, h - 1), np.minimum(ix, w - 1)]


INTERP = {"nearest": interp_nearest, "linear": interp_linear, "poly3": interp_poly3,
          "poly5": interp_poly5}


def blot_phases(src, coef, ny, nx, phases=((0.0, 0.0),), scale=1.0, misval=0.0,
                interp="poly5"):
    """Blot float32 ``src`` onto a (ny, nx) target grid for every phase
    displacement ``(ddx, ddy)`` (target pixel (i, j) samples the map at
    (i + ddx, j + ddy)); returns float32 [nph, ny, nx]."""
    src = np.asarray(src, F)
    h, w = src.shape
    jj, ii = np.meshgrid(np.arange(ny, dtype=np.float64), np.arange(nx, dtype=np.float64),
                         indexing="ij")
    out = np.empty((len(phases), ny, nx), F)
    fn = INTERP[interp]
    for k, (ddx, ddy) in enumerate(phases):
        xs, ys = eval_map(coef, ii + float(ddx), jj + float(ddy))
        xf, yf = xs.astype(F), ys.astype(F)
        # blot.py:118-121 + doblot: 1-based xmin=1 <= x+1 <= xmax=naxis1
        ok = (xf >= 0) & (xf <= F(w - 1)) & (yf >= 0) & (yf <= F(h - 1))
        xq = np.where(ok, xf, F(0))
        yq = np.where(ok, yf, F(0))
        v = fn(src, xq, yq) * F(scale)
        out[k] = np.where(ok, v, F(misval))
    return out
