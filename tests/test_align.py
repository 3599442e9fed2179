"""Batched caller: host-side linear fit on the CPU,
and the reference-signature find_linear_fit on the GPU with a stand-in blot."""
import copy
import warnings

import numpy as np
import pytest

from oracle import subpixal_oracle as orc


def test_iter_linear_fit_recovers_affine_with_outliers():
    from subpixal_b200.align import iter_linear_fit
    rng = np.random.default_rng(1)
    xy = rng.uniform(0, 4096, (400, 2))
    th = np.deg2rad(0.02)
    m = 1.00003 * np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
    c = np.array([2048.0, 2048.0])
    uv = (xy - c) @ m.T + c + np.array([0.31, -0.17]) + 0.002 * rng.standard_normal((400, 2))
    uv[::37] += 5.0                                    # outliers
    for geom in ("rscale", "general"):
        fit = iter_linear_fit(xy, uv, fitgeom=geom, center=c, nclip=3, sigma=3.0)
        assert np.allclose(fit["offset"], [0.31, -0.17], atol=2e-3)
        assert np.allclose(fit["matrix"], m, atol=2e-6)
        assert abs(fit["rot"] - 0.02) < 1e-3 and abs(fit["scale"][0] - 1.00003) < 2e-6
        assert fit["proper"] and fit["fitmask"].sum() <= 400 - 11 and fit["rmse"] < 0.01
        assert fit["resids"].shape == (fit["fitmask"].sum(), 2)
    fit = iter_linear_fit(xy, xy + [0.5, -0.25], fitgeom="shift")
    assert np.allclose(fit["offset"], [0.5, -0.25]) and fit["rot"] == 0.0
    w = np.ones(400); w[::37] = 0.0
    fit = iter_linear_fit(xy, uv, wuv=w, fitgeom="general", center=c, nclip=0)
    assert np.allclose(fit["offset"], [0.31, -0.17], atol=2e-3)
    with pytest.raises(ValueError):
        iter_linear_fit(xy, uv, fitgeom="affine")
    with pytest.raises(ValueError):
        iter_linear_fit(xy[:2], uv[:2], fitgeom="general")


class _Wcs:
    """Identity tangent plane: world == 0-based pixel coordinates."""
    sip = cpdis1 = cpdis2 = det2im1 = det2im2 = None

    class wcs:
        crpix = [10.0, 20.0]

    def deepcopy(self):
        return copy.deepcopy(self)

    def all_pix2world(self, x, y, origin):
        return np.asarray(x, float) - origin, np.asarray(y, float) - origin

    def all_world2pix(self, ra, dec, origin, **kw):
        return np.asarray(ra, float) + origin, np.asarray(dec, float) + origin

    def wcs_world2pix(self, ra, dec, origin):
        return np.asarray(ra, float) + origin, np.asarray(dec, float) + origin


@pytest.mark.gpu
def test_find_linear_fit_matches_per_source_loop():
    from subpixal_b200 import align, cutout
    rng = np.random.default_rng(3)
    sizes = [(32, 32)] * 6 + [(24, 40)] * 5 + [(64, 64)] * 4
    stacks = [orc.random_pairs(1, ny, nx, seed=100 + k)[0][0] for k, (ny, nx) in enumerate(sizes)]
    frame = np.zeros((400, 400), dtype=np.float32)
    imgs, drzs, phases = [], [], {}
    for k, ((ny, nx), st) in enumerate(zip(sizes, stacks)):
        x1, y1 = 5 + 20 * k, 7 + 15 * k
        frame[y1:y1 + ny, x1:x1 + nx] = st[0]
        im = cutout.Cutout(frame, _Wcs(), blc=(x1, y1), trc=(x1 + nx - 1, y1 + ny - 1),
                           src_id=k, src_weight=1.0 + k)
        dz = cutout.Cutout(frame, _Wcs(), blc=(x1, y1), trc=(x1 + nx - 1, y1 + ny - 1),
                           src_id=k, src_weight=1.0 + k)
        assert np.array_equal(im.data, st[0])
        imgs.append(im); drzs.append(dz)
        phases[k] = st

    def blot(dzct, imct):          # stand-in for subpixal.blot.blot_cutout
        out = copy.deepcopy(imct)
        ph = {(0, 0): 1, (-0.5, 0): 2, (0, -0.5): 3, (-0.5, -0.5): 4}[(imct.dx, imct.dy)]
        out.data = phases[dzct.src_id][ph].copy()
        return out

    fit, iccs, blts = align.find_linear_fit(imgs, drzs, fitgeom="general", blot=blot)
    assert len(iccs) == len(blts) == len(sizes) and all(c.dx == 0 and c.dy == 0 for c in imgs)
    want = []
    for st in stacks:
        with warnings.catch_warnings(), np.errstate(all="ignore"):
            warnings.simplefilter("ignore")
            want.append(orc.find_displacement(*st, cc_type="NCC", full_output=True))
    for k, w in enumerate(want):
        assert abs(fit["subpixal_img_dxy"][k, 0] - w[0]) < 1e-4
        assert abs(fit["subpixal_img_dxy"][k, 1] - w[1]) < 1e-4
        assert iccs[k].shape == w[2].shape
        assert np.linalg.norm(iccs[k] - w[2]) / np.linalg.norm(w[2]) < 1e-5
    assert np.allclose(fit["subpixal_ref_dxy"], -fit["subpixal_img_dxy"], atol=1e-9)
    ref = align.iter_linear_fit(
        np.array([np.add(c.cutout_src_pos, c.blc) + 1 for c in imgs]),
        np.array([np.add(c.cutout_src_pos, c.blc) + 1 for c in imgs]) + fit["subpixal_img_dxy"],
        wuv=np.array([c.src_weight for c in drzs]), fitgeom="general", center=[10.0, 20.0])
    assert np.allclose(ref["offset"], fit["offset"], atol=1e-9)
    assert {"offset", "rot", "rotxy", "scale", "proper", "skew", "matrix", "resids", "rmse",
            "mae", "fitmask", "irmse"} <= set(fit)
    with pytest.raises(ValueError):
        align.find_linear_fit(imgs, drzs[:-1], blot=blot)
