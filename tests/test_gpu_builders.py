"""GPU tests of the cutout list builders and the segmentation bounding-box
kernel against loops written with the CPU oracle (restating reference
cutout.py:47-587 on top of oracle.crop / oracle.bitfield_bad_mask)."""
import numpy as np
import pytest
import torch

from oracle import subpixal_oracle as orc

pytestmark = pytest.mark.gpu


class AffineWCS:
    """world = A @ pix + b (0-based pixels), astropy-like calling conventions."""
    def __init__(self, a, b):
        self.a, self.b = np.asarray(a, float), np.asarray(b, float)
        self.ai = np.linalg.inv(self.a)
        self.pixel_scale_matrix = self.a

    @staticmethod
    def _args(args):
        if len(args) == 2:                       # (Nx2 array, origin)
            arr = np.asarray(args[0], float)
            return arr[:, 0], arr[:, 1], args[1], True
        return np.asarray(args[0], float), np.asarray(args[1], float), args[2], False

    def all_pix2world(self, *args, **kw):
        x, y, o, packed = self._args(args)
        w = np.stack([x - o, y - o], 1) @ self.a.T + self.b
        return w if packed else (w[:, 0], w[:, 1])

    def all_world2pix(self, *args, **kw):
        r, d, o, packed = self._args(args)
        p = (np.stack([r, d], 1) - self.b) @ self.ai.T + o
        return p if packed else (p[:, 0], p[:, 1])


def _scene(seed=0, ny=120, nx=150, nsrc=14):
    rng = np.random.default_rng(seed)
    img = rng.standard_normal((ny, nx)).astype(np.float32)
    seg = np.zeros((ny, nx), dtype=np.int32)
    ids, xs, ys = [], [], []
    for k in range(nsrc):
        sid = 3 * k + 2
        w, h = rng.integers(3, 14, 2)
        x1, y1 = rng.integers(0, nx - w), rng.integers(0, ny - h)
        blob = rng.uniform(size=(h, w)) < 0.8
        seg[y1:y1 + h, x1:x1 + w][blob] = sid
        ids.append(sid); xs.append(x1 + w / 2.0); ys.append(y1 + h / 2.0)
    seg[0, 5] = 999                               # touches the edge -> skipped
    ids += [999, 12345]                            # one id without a segment
    xs += [5.0, 1.0]; ys += [0.0, 1.0]
    img[10, 10] = np.nan
    dq = rng.integers(0, 8, (ny, nx)).astype(np.uint16)
    dq[rng.uniform(size=(ny, nx)) < 0.7] = 0
    cat = {"id": np.array(ids), "x": np.array(xs), "y": np.array(ys),
           "weight": np.linspace(1.0, 2.0, len(ids))}
    return img, seg, dq, cat


def test_segment_boxes_kernel():
    from subpixal_b200 import batch
    rng = np.random.default_rng(4)
    seg = rng.integers(0, 60, (97, 131)).astype(np.int32)
    seg[rng.uniform(size=seg.shape) < 0.6] = 0
    seg[50:60, 20:100] = 77
    for dt in (np.int32, np.int64):
        tab = batch.segment_boxes(torch.from_numpy(seg.astype(dt)).cuda()).cpu().numpy()
        assert tab.shape == (78, 4)
        for lab in range(1, 78):
            yy, xx = np.nonzero(seg == lab)
            if yy.size == 0:
                assert tab[lab, 1] == -1 and tab[lab, 3] == -1
            else:
                assert list(tab[lab]) == [xx.min(), xx.max(), yy.min(), yy.max()]
    empty = batch.segment_boxes(torch.zeros((5, 7), dtype=torch.int32, device="cuda"))
    assert empty.shape == (1, 4)


def _primary_reference_loop(cat, seg, img, dq, dqbitmask, pad):
    """cutout.py:139-196 with the oracle's crop."""
    ny, nx = seg.shape
    ids, cidx, _ = np.intersect1d(cat["id"], np.setdiff1d(np.unique(seg), [0]),
                                  return_indices=True)
    out = []
    for sid, ci in zip(ids, cidx):
        m = seg == sid
        yy, xx = np.where(m)
        x1, x2, y1, y2 = xx.min(), xx.max(), yy.min(), yy.max()
        if x1 <= 0 or y1 <= 0 or x2 >= nx - 1 or y2 >= ny - 1:
            continue
        blc, trc = (x1 - pad, y1 - pad), (x2 + pad, y2 + pad)
        c = orc.crop(img, blc, trc, dq=dq, mode="strict", fillval=0)
        mask = c["mask"] | ~m[c["extraction_slice"]]
        mask |= orc.bitfield_bad_mask(c["dq"], dqbitmask)
        if not np.all(mask):
            out.append((sid, blc, trc, c["data"], mask, c["dq"],
                        (cat["x"][ci], cat["y"][ci]), cat["weight"][ci]))
    return out


@pytest.mark.parametrize("dqbitmask", [0, 6, None, "~1"])
def test_create_primary_cutouts(dqbitmask):
    from subpixal_b200 import cutout
    img, seg, dq, cat = _scene(seed=1)
    wcs = AffineWCS(np.eye(2), [0, 0])
    got = cutout.create_primary_cutouts(cat, seg, img, wcs, imdq=dq, dqbitmask=dqbitmask,
                                        data_units="rate", exptime=3.0, pad=1)
    want = _primary_reference_loop(cat, seg, img, dq, cutout.interpret_bit_flags(dqbitmask), 1)
    assert len(got) == len(want) > 5
    for g, (sid, blc, trc, data, mask, cdq, pos, wgt) in zip(got, want):
        assert g.src_id == sid and g.blc == blc and g.trc == trc
        assert np.array_equal(g.data, data, equal_nan=True) and np.array_equal(g.mask, mask)
        assert np.array_equal(g.dq, cdq) and g.dq.dtype == dq.dtype
        assert g.src_pos == pos and g.src_weight == wgt and g.exptime == 3.0
        assert g.wcs is wcs and g.data_units == "rate"


def test_create_cutouts_roundtrip():
    """primary (drz) -> flt -> drz with an affine 'distortion' between the images."""
    from subpixal_b200 import cutout
    drz, seg, dq, cat = _scene(seed=2, ny=160, nx=180, nsrc=10)
    drz = np.nan_to_num(drz)
    rng = np.random.default_rng(9)
    flt = rng.standard_normal((150, 170)).astype(np.float32)
    drz_wcs = AffineWCS(np.eye(2) * 0.9, [100.0, 50.0])
    th = 0.05
    flt_wcs = AffineWCS([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]], [102.0, 48.0])
    prim = cutout.create_primary_cutouts(cat, seg, drz, drz_wcs, pad=1)
    flt_cts, drz_cts = cutout.create_cutouts(prim, seg, drz, drz_wcs, flt, flt_wcs, pad=2)
    assert len(flt_cts) == len(drz_cts) > 3
    for f, d in zip(flt_cts, drz_cts):
        assert f.src_id == d.src_id
        # the flt cutout is the oracle crop of its own box
        o = orc.crop(flt, f.blc, f.trc, mode="strict", fillval=0)
        assert np.array_equal(f.data, o["data"]) and np.array_equal(f.mask, o["mask"])
        # the drz cutout: 'fill' crop + segmentation mask (cutout.py:416-419)
        o = orc.crop(drz, d.blc, d.trc, mode="fill", fillval=0)
        sg = np.zeros(o["data"].shape, dtype=seg.dtype)
        sg[o["insertion_slice"]] = seg[o["extraction_slice"]]
        assert np.array_equal(d.data, o["data"])
        assert np.array_equal(d.mask, o["mask"] | (sg != d.src_id))
        # source positions map to the same sky point (cutout.py:300-301)
        wf = f.pix2world([list(f.cutout_src_pos)])
        wd = d.pix2world([list(d.cutout_src_pos)])
        assert np.allclose(wf, wd, atol=1e-9)
        # the flt footprint encloses the primary cutout's footprint (+pad)
        fp = flt_wcs.all_world2pix(d.get_bbox("world"), 0)
        assert f.blc[0] - 8 <= fp[:, 0].min() and fp[:, 0].max() <= f.trc[0] + 8
    assert abs(flt_cts[0].pscale - 1.0) < 1e-12 and abs(drz_cts[0].pscale - 0.9) < 1e-12
