"""Blot row (reference blot.py:79-155, align.py:661-679).

CPU tier: properties of the oracle restatement (parity with drizzlepac is UNPINNED,
see oracle/blot_oracle.py) and the host logic of ``subpixal_b200.blot``.
GPU tier: the CUDA kernel through the C-ABI, bit-for-bit against the oracle, and
the batched ``find_linear_fit`` with the on-device blot."""
import copy

import numpy as np
import pytest
import torch

from oracle import blot_oracle as bo

PHASES = ((0.0, 0.0), (0.5, 0.0), (0.0, 0.5), (0.5, 0.5))


def _poly(x, y):
    return (1 + 0.3 * x - 0.2 * y + 0.01 * x * y + 1e-3 * x ** 2 - 2e-3 * y ** 2
            + 1e-5 * x ** 3 * y - 3e-6 * x ** 2 * y ** 3 + 1e-7 * x ** 5)


# ------------------------------------------------------------------ CPU: oracle
def test_oracle_poly5_reproduces_quintics_in_the_interior():
    h, w = 40, 50
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    data = _poly(xx, yy).astype(np.float32)
    rng = np.random.default_rng(0)
    x = rng.uniform(2, w - 4, 2000).astype(np.float32)
    y = rng.uniform(2, h - 4, 2000).astype(np.float32)
    want = _poly(x.astype(np.float64), y.astype(np.float64))
    got = bo.interp_poly5(data, x, y)
    assert got.dtype == np.float32
    assert np.max(np.abs(got - want)) < 4e-6 * np.max(np.abs(want))      # fp32 rounding only


def test_oracle_edge_extension_is_exact_for_planes_and_hits_the_samples():
    h, w = 17, 23
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    plane = (2 + 0.5 * xx - 0.25 * yy).astype(np.float32)
    rng = np.random.default_rng(1)
    x = rng.uniform(0, w - 1, 3000).astype(np.float32)
    y = rng.uniform(0, h - 1, 3000).astype(np.float32)
    got = bo.interp_poly5(plane, x, y)
    assert np.max(np.abs(got - (2 + 0.5 * x.astype(np.float64) - 0.25 * y))) < 1e-5
    data = rng.standard_normal((h, w)).astype(np.float32)
    jj, ii = np.mgrid[0:h, 0:w]
    cubic = (1 + 0.3 * xx - 0.2 * yy + 0.01 * xx * yy + 1e-3 * xx ** 2 - 2e-4 * yy ** 3).astype(np.float32)
    xi = rng.uniform(1, w - 3, 500).astype(np.float32); yi = rng.uniform(1, h - 3, 500).astype(np.float32)
    want = (1 + 0.3 * xi.astype(np.float64) - 0.2 * yi + 0.01 * xi.astype(np.float64) * yi
            + 1e-3 * xi.astype(np.float64) ** 2 - 2e-4 * yi.astype(np.float64) ** 3)
    assert np.max(np.abs(bo.interp_poly3(cubic, xi, yi) - want)) < 2e-5     # cubics are reproduced
    for fn in (bo.interp_poly5, bo.interp_poly3, bo.interp_linear, bo.interp_nearest):
        got = fn(data, ii.ravel().astype(np.float32), jj.ravel().astype(np.float32))
        assert np.allclose(got.reshape(h, w), data, atol=2e-6)
    # a 6x6 stencil on an image smaller than the stencil still works (folded + clamped)
    tiny = rng.standard_normal((3, 4)).astype(np.float32)
    assert np.all(np.isfinite(bo.interp_poly5(tiny, np.float32([0.5, 2.9]), np.float32([1.2, 0.1]))))


def test_oracle_blot_identity_shift_and_misval():
    h, w = 30, 36
    rng = np.random.default_rng(2)
    data = rng.standard_normal((h, w)).astype(np.float32)
    ident = bo.fit_map(lambda x, y: (x, y), w, h)
    out = bo.blot_phases(data, ident, h, w)
    assert np.allclose(out[0], data, atol=2e-6)
    # the half-pixel phases of a smooth image are the image sampled half a pixel further on
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    smooth = (100 * np.exp(-((xx - 17.3) ** 2 + (yy - 14.1) ** 2) / 30.0)).astype(np.float32)
    ph = bo.blot_phases(smooth, ident, h, w, phases=PHASES)
    want = 100 * np.exp(-((xx + 0.5 - 17.3) ** 2 + (yy - 14.1) ** 2) / 30.0)
    assert np.max(np.abs(ph[1][:, :-1] - want[:, :-1])) < 5e-3          # quintic on a sigma=3.9 px Gaussian
    assert np.all(ph[1][:, -1] == 0) and np.all(ph[2][-1, :] == 0)      # mapped outside -> misval
    # flux scale and a custom misval
    out = bo.blot_phases(smooth, bo.fit_map(lambda x, y: (x + 20.0, y), w, h), h, w,
                         scale=0.25, misval=-7.0)
    assert np.all(out[0][:, 16:] == -7.0)
    assert np.allclose(out[0][:, :15], 0.25 * smooth[:, 20:35], atol=1e-6)


def test_product_map_fit_matches_the_oracle_and_is_exact_for_biquadratics():
    from subpixal_b200 import blot
    f = lambda x, y: (1.01 * x + 0.3 + 1e-5 * x * y - 3e-6 * y * y, 0.99 * y - 0.2 + 2e-5 * x * x)
    one_based = lambda x, y: tuple(v + 1.0 for v in f(np.asarray(x) - 1.0, np.asarray(y) - 1.0))
    got = blot.fit_pixel_map(one_based, 50, 40)
    want = bo.fit_map(f, 50, 40)
    assert np.allclose(got, want, rtol=0, atol=1e-12)
    x = np.array([0.0, 3.0, 17.5, 49.0]); y = np.array([39.0, 5.0, 30.25, 0.0])
    xs, ys = bo.eval_map(got, x, y)
    assert np.allclose(xs, f(x, y)[0], atol=1e-11) and np.allclose(ys, f(x, y)[1], atol=1e-11)


class _AffineWcs:
    """Tangent plane with world = M @ pix0 + t (0-based pixels); no distortion."""
    sip = cpdis1 = cpdis2 = det2im1 = det2im2 = None

    def __init__(self, m=((1.0, 0.0), (0.0, 1.0)), t=(0.0, 0.0), crpix=(10.0, 20.0)):
        self.m = np.asarray(m, float)
        self.t = np.asarray(t, float)
        self.pixel_scale_matrix = self.m

        class _W:
            pass
        self.wcs = _W()
        self.wcs.crpix = list(crpix)

    def deepcopy(self):
        return copy.deepcopy(self)

    def all_pix2world(self, x, y, origin):
        p = np.stack([np.asarray(x, float) - origin, np.asarray(y, float) - origin])
        w = self.m @ p.reshape(2, -1) + self.t[:, None]
        return w[0].reshape(np.shape(x)), w[1].reshape(np.shape(x))

    def all_world2pix(self, ra, dec, origin, **kw):
        w = np.stack([np.asarray(ra, float), np.asarray(dec, float)]).reshape(2, -1)
        p = np.linalg.solve(self.m, w - self.t[:, None]) + origin
        return p[0].reshape(np.shape(ra)), p[1].reshape(np.shape(ra))

    wcs_world2pix = all_world2pix


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only behaviour")
def test_blot_has_no_cpu_fallback_and_checks_arguments():
    from subpixal_b200 import blot, cutout
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        blot.blot_phases([], [])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        cutout.Cutout(np.zeros((8, 8), np.float32), _AffineWcs(), blc=(0, 0), trc=(3, 3))


def test_blot_interp_names():
    from subpixal_b200 import blot

    class _C:
        naxis = [4, 4]
    with pytest.raises(NotImplementedError):
        blot._check(_C(), _C(), "sinc")
    with pytest.raises(ValueError):
        blot._check(_C(), _C(), "cubic")

    class _C3:
        naxis = [4, 4, 2]
    with pytest.raises(ValueError, match="2D"):
        blot._check(_C3(), _C(), "poly5")


# ------------------------------------------------------------------ GPU: kernel
def _random_case(rng, k):
    h, w = (int(v) for v in rng.integers(4, 90, 2))
    ny, nx = (int(v) for v in rng.integers(3, 80, 2))
    src = rng.standard_normal((h, w)).astype(np.float32) * 10 + 3
    sc = w / nx * rng.uniform(0.6, 1.3)
    th = rng.uniform(-0.2, 0.2)
    a = sc * np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
    t = rng.uniform(-6, 6, 2)
    q = rng.uniform(-2e-4, 2e-4, 4)
    f = lambda x, y: (a[0, 0] * x + a[0, 1] * y + t[0] + q[0] * x * y + q[1] * y * y,
                      a[1, 0] * x + a[1, 1] * y + t[1] + q[2] * x * x + q[3] * x * y)
    return src, bo.fit_map(f, nx, ny), (ny, nx)


@pytest.mark.gpu
@pytest.mark.parametrize("interp,prepare", [("poly5", True), ("poly5", False), ("poly3", False),
                                            ("linear", False), ("nearest", False)])
def test_gpu_blot_is_bit_identical_to_the_oracle(interp, prepare):
    from subpixal_b200 import batch
    rng = np.random.default_rng(7)
    cases = [_random_case(rng, k) for k in range(24)]
    # identity maps: every pixel sits exactly on a sample / on the upper edge
    for (h, w) in ((6, 6), (33, 20), (64, 64)):
        src = rng.standard_normal((h, w)).astype(np.float32)
        cases.append((src, bo.fit_map(lambda x, y: (x, y), w, h), (h, w)))
    flat = np.concatenate([c[0].ravel() for c in cases])
    s_shp = np.array([c[0].shape for c in cases])
    s_off = np.concatenate([[0], np.cumsum(s_shp[:, 0] * s_shp[:, 1])])[:-1]
    maps = np.stack([c[1] for c in cases])
    t_shp = np.array([c[2] for c in cases])
    scale = rng.uniform(0.5, 2.0, len(cases)).astype(np.float32)
    out, off = batch.blot_phases(torch.from_numpy(flat).cuda(), s_off, s_shp, maps, t_shp,
                                 phases=PHASES, scale=scale, interp=interp, misval=-1.5,
                                 prepare=prepare)
    out = out.cpu().numpy()
    inside = 0
    for k, (src, coef, (ny, nx)) in enumerate(cases):
        want = bo.blot_phases(src, coef, ny, nx, phases=PHASES, scale=scale[k], misval=-1.5,
                              interp=interp)
        got = out[off[k]:off[k + 1]].reshape(4, ny, nx)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (k, interp)
        inside += int(np.sum(want != -1.5))
    assert inside > 10000


@pytest.mark.gpu
def test_gpu_blot_cutout_and_linear_fit_with_device_blot():
    """drz frame on the identity WCS, image frame rendered through a true affine WCS but
    carrying a WCS that is off by a known shift: blot + CC must measure that shift."""
    from subpixal_b200 import align, blot, cutout
    rng = np.random.default_rng(11)
    H = W = 420
    nsrc = 30
    cx = rng.uniform(40, W - 40, nsrc); cy = rng.uniform(40, H - 40, nsrc)
    amp = rng.uniform(200, 2000, nsrc); sig = rng.uniform(1.8, 3.0, nsrc)

    def scene(wx, wy):
        v = np.zeros(np.shape(wx))
        for k in range(nsrc):
            v += amp[k] * np.exp(-((wx - cx[k]) ** 2 + (wy - cy[k]) ** 2) / (2 * sig[k] ** 2))
        return v

    yy, xx = np.mgrid[0:H, 0:W].astype(np.float64)
    drz = scene(xx, yy).astype(np.float32)
    th = 2e-3
    m = 1.002 * np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
    t_true = np.array([3.2, -2.6])
    err = np.array([0.31, -0.17])                       # image pixels
    wtrue = _AffineWcs(m, t_true)
    wx, wy = wtrue.all_pix2world(xx, yy, 0)
    img = (scene(wx, wy) + rng.standard_normal((H, W))).astype(np.float32)
    # the image's catalogued WCS: pixel p is believed to see world(p + err)
    wbad = _AffineWcs(m, t_true + m @ err)
    wdrz = _AffineWcs()
    imgs, drzs = [], []
    half = 16
    for k in range(nsrc):
        px, py = np.linalg.solve(m, np.array([cx[k], cy[k]]) - t_true)
        x1, y1 = int(round(px)) - half, int(round(py)) - half
        if x1 < 8 or y1 < 8 or x1 + 2 * half + 8 > W or y1 + 2 * half + 8 > H:
            continue
        imgs.append(cutout.Cutout(img, wbad, blc=(x1, y1), trc=(x1 + 2 * half - 1, y1 + 2 * half - 1),
                                  src_id=k, src_weight=1.0))
        dx1, dy1 = int(round(cx[k])) - 24, int(round(cy[k])) - 24
        drzs.append(cutout.Cutout(drz, wdrz, blc=(dx1, dy1), trc=(dx1 + 47, dy1 + 47),
                                  src_id=k, src_weight=1.0, mode='fill', fillval=0.0))
    assert len(imgs) >= 15
    # single blot through the reference-signature wrapper == oracle on the same map
    b = blot.blot_cutout(drzs[0], imgs[0])
    assert type(b) is type(imgs[0]) and b.data.shape == imgs[0].data.shape and b is not imgs[0]
    coef = blot.fit_pixel_map(blot.BlotWCSMap(imgs[0], drzs[0]), 2 * half, 2 * half)
    pr = blot.pixel_ratio(drzs[0], imgs[0])
    want = bo.blot_phases(drzs[0].data, coef, 2 * half, 2 * half, scale=np.float32(1.0 / pr ** 2))[0]
    assert np.array_equal(b.data, want)
    with pytest.raises(NotImplementedError):
        blot.blot_cutout(drzs[0], imgs[0], interp='lanczos3')
    fit, iccs, blts = align.find_linear_fit(imgs, drzs, fitgeom='shift', nclip=0)
    d = fit['subpixal_img_dxy']
    assert len(iccs) == len(blts) == len(imgs) and all(c.dx == 0 and c.dy == 0 for c in imgs)
    # the blot lands `err` pixels away from where the image really is
    assert np.max(np.abs(np.median(d, axis=0) - (-err))) < 0.02 or \
        np.max(np.abs(np.median(d, axis=0) - err)) < 0.02
    assert np.array_equal(blts[0].data, b.data)


@pytest.mark.gpu
def test_frame_pipeline_matches_the_cutout_object_path():
    """align.frame_displacements (gather + gather + blot + ragged CC, all on the
    device) against find_linear_fit over Cutout objects on the same frames."""
    from subpixal_b200 import align, blot, cutout
    rng = np.random.default_rng(5)
    H = W = 360
    nsrc = 24
    cx = rng.uniform(45, W - 45, nsrc); cy = rng.uniform(45, H - 45, nsrc)
    amp = rng.uniform(300, 3000, nsrc); sig = rng.uniform(1.6, 3.2, nsrc)
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float64)

    def scene(wx, wy):
        v = np.zeros(np.shape(wx))
        for k in range(nsrc):
            v += amp[k] * np.exp(-((wx - cx[k]) ** 2 + (wy - cy[k]) ** 2) / (2 * sig[k] ** 2))
        return v

    drz = scene(xx, yy).astype(np.float32)
    th = -1.5e-3
    m = 0.998 * np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
    t = np.array([-1.7, 2.4])
    wimg = _AffineWcs(m, t)
    wx, wy = wimg.all_pix2world(xx, yy, 0)
    img = (scene(wx + 0.2, wy - 0.1) + rng.standard_normal((H, W))).astype(np.float32)
    wdrz = _AffineWcs()
    ib, db, imgs, drzs = [], [], [], []
    for k in range(nsrc):
        px, py = np.linalg.solve(m, np.array([cx[k], cy[k]]) - t)
        hw, hh = int(rng.integers(10, 30)), int(rng.integers(10, 30))       # ragged boxes
        x1, y1 = int(round(px)) - hw, int(round(py)) - hh
        X1, Y1 = int(round(cx[k])) - hw - 8, int(round(cy[k])) - hh - 8
        ib.append((x1, y1, x1 + 2 * hw - 1, y1 + 2 * hh - 1))
        db.append((X1, Y1, X1 + 2 * hw + 15, Y1 + 2 * hh + 15))
        imgs.append(cutout.Cutout(img, wimg, blc=ib[-1][:2], trc=ib[-1][2:], src_id=k, mode='fill',
                                  fillval=0.0))
        drzs.append(cutout.Cutout(drz, wdrz, blc=db[-1][:2], trc=db[-1][2:], src_id=k, mode='fill',
                                  fillval=0.0))
    fit, iccs, _ = align.find_linear_fit(imgs, drzs, fitgeom='shift', nclip=0, use_weights=False)
    want = fit['subpixal_img_dxy']

    def frame_map(x, y):
        return wdrz.all_world2pix(*wimg.all_pix2world(x, y, 0), 0)

    pr = blot.pixel_ratio(drzs[0], imgs[0])
    sh, pk, st, icc, ioff = align.frame_displacements(
        torch.from_numpy(img).cuda(), torch.from_numpy(drz).cuda(), ib, db, frame_map,
        flux_scale=1.0 / pr ** 2, full_output=True)
    got = sh.cpu().numpy()
    assert np.max(np.abs(got - want)) < 1e-5
    assert np.max(np.abs(np.median(got, axis=0) - np.array([0.2, -0.1]) / 0.998)) < 0.03 or \
        np.max(np.abs(np.median(got, axis=0) + np.array([0.2, -0.1]) / 0.998)) < 0.03
    k = 3
    ny, nx = ib[k][3] - ib[k][1] + 1, ib[k][2] - ib[k][0] + 1
    g = icc.cpu().numpy()[ioff[k]:ioff[k + 1]].reshape(2 * ny, 2 * nx)
    assert np.linalg.norm(g - iccs[k]) / np.linalg.norm(iccs[k]) < 1e-5
    assert int((st.cpu().numpy() != 0).sum()) == 0
    # chunked processing (device scratch budget) gives the same table and the same surfaces
    sh2, pk2, st2, icc2, ioff2 = align.frame_displacements(
        torch.from_numpy(img).cuda(), torch.from_numpy(drz).cuda(), ib, db, frame_map,
        flux_scale=1.0 / pr ** 2, full_output=True, max_scratch_bytes=150000)
    assert torch.equal(sh, sh2) and torch.equal(pk, pk2) and torch.equal(st, st2)
    assert torch.equal(icc, icc2) and np.array_equal(np.asarray(ioff), np.asarray(ioff2))


def test_oracle_matches_its_frozen_outputs():
    """tests/golden/blot_oracle.npz (made by tests/golden/make_blot_golden.py) freezes the
    restatement - it is NOT a reference output, drizzlepac being unavailable."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "blot_oracle.npz"))
    for k in range(4):
        ny, nx = (int(v) for v in g["shape%d" % k])
        for interp in ("poly5", "poly3", "linear", "nearest"):
            got = bo.blot_phases(g["src%d" % k], g["map%d" % k], ny, nx, phases=PHASES, scale=0.75,
                                 misval=-2.0, interp=interp)
            assert np.array_equal(got.view(np.uint32), g["%s%d" % (interp, k)].view(np.uint32)), (k, interp)


@pytest.mark.gpu
def test_gpu_blot_matches_the_frozen_oracle_outputs():
    import os
    from subpixal_b200 import batch
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "blot_oracle.npz"))
    srcs = [g["src%d" % k] for k in range(4)]
    shp = np.array([s.shape for s in srcs])
    off = np.concatenate([[0], np.cumsum(shp[:, 0] * shp[:, 1])])[:-1]
    tg = np.array([g["shape%d" % k] for k in range(4)])
    maps = np.stack([g["map%d" % k] for k in range(4)])
    flat = torch.from_numpy(np.concatenate([s.ravel() for s in srcs])).cuda()
    for interp in ("poly5", "poly3", "linear", "nearest"):
        out, o = batch.blot_phases(flat, off, shp, maps, tg, phases=PHASES, scale=np.full(4, 0.75, np.float32),
                                   interp=interp, misval=-2.0)
        out = out.cpu().numpy()
        for k in range(4):
            want = g["%s%d" % (interp, k)]
            assert np.array_equal(out[o[k]:o[k + 1]].reshape(want.shape).view(np.uint32), want.view(np.uint32))


def test_everett_form_equals_six_point_lagrange():
    """A structurally independent witness for the interpolator's algebra (the blot row
    stays PARITY UNPINNED - drizzlepac is absent - but the Everett central-difference
    form the oracle and the kernel share must at least be THE quintic through the six
    samples): compare it with the float64 Lagrange polynomial through the same six
    points, in 1-D and as the 6x6 tensor product of `interp_poly5`."""
    from oracle import blot_oracle as bo
    rng = np.random.default_rng(12)

    def lagrange6(f, s):                      # nodes -2..3, evaluated at s, float64
        nodes = np.arange(-2, 4, dtype=np.float64)
        out = 0.0
        for k in range(6):
            w = 1.0
            for m in range(6):
                if m != k:
                    w *= (s - nodes[m]) / (nodes[k] - nodes[m])
            out += w * float(f[k])
        return out

    for _ in range(200):
        f = rng.standard_normal(6).astype(np.float32) * 10
        s = np.float32(rng.uniform(0, 1))
        got = float(bo._everett(*f, s))
        want = lagrange6(f.astype(np.float64), float(s))
        assert abs(got - want) < 1e-5 * max(1.0, np.abs(f).max())
    # 2-D: interior points of a random image (no edge extension involved)
    img = rng.standard_normal((24, 24)).astype(np.float32)
    for _ in range(100):
        x, y = rng.uniform(4, 18, 2)
        got = float(bo.interp_poly5(img, np.float32(x), np.float32(y)))
        ix, iy = int(x), int(y)
        sx, sy = float(np.float32(x)) - ix, float(np.float32(y)) - iy
        rows = [lagrange6(img[iy + dj, ix - 2:ix + 4].astype(np.float64), sx) for dj in range(-2, 4)]
        want = lagrange6(np.array(rows), sy)
        assert abs(got - want) < 2e-5 * max(1.0, float(np.abs(img).max()))
