"""GPU tests of the round-1 advisor findings: FITS byte order, integer frames,
oversize pairs in ragged batches, the DQ plane of the device pipeline."""
import warnings

import numpy as np
import pytest
import torch

from oracle import subpixal_oracle as orc

pytestmark = pytest.mark.gpu


def test_big_endian_and_integer_frames():
    """astropy.io.fits hands out '>f4' / '>i2' arrays; the reference crops any dtype
    (cutout.py:737-752)."""
    from subpixal_b200 import cutout, centroid
    rng = np.random.default_rng(5)
    img = rng.standard_normal((40, 50)).astype(np.float32)
    img[3, 4] = np.nan
    dq = rng.integers(0, 4, (40, 50)).astype(np.int16)
    wht = rng.uniform(size=(40, 50)).astype(np.float32)
    want = orc.crop(img, (-3, 2), (20, 30), dq=dq, weight=wht, mode="fill", fillval=0)
    got = cutout.Cutout(img.astype(">f4"), None, blc=(-3, 2), trc=(20, 30), dq=dq.astype(">i2"),
                        weight=wht.astype(">f4"), mode="fill", fillval=0)
    assert np.array_equal(got.data, want["data"], equal_nan=True)
    assert np.array_equal(got.mask, want["mask"])
    assert np.array_equal(got.dq, want["dq"]) and np.array_equal(got.weight, want["weight"])
    assert got.data.dtype == np.dtype(">f4")               # dtype of the input, like numpy
    # integer frames: exact crop, fill value in the frame's dtype, nothing non-finite
    for dt in (np.int16, np.uint16, np.int32, np.int64, np.uint8):
        ints = rng.integers(0, 200, (40, 50)).astype(dt)
        want = orc.crop(ints, (-3, 2), (20, 30), mode="fill", fillval=7)
        got = cutout.Cutout(ints, None, blc=(-3, 2), trc=(20, 30), mode="fill", fillval=7)
        assert got.data.dtype == dt and np.array_equal(got.data, want["data"])
        assert np.array_equal(got.mask, want["mask"])
    # the default fillval (NaN) in an integer frame is whatever np.full makes of it
    # (cutout.py:737; a RuntimeWarning and INT_MIN with numpy 2.x) - same call here
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want = orc.crop(ints, (-3, 2), (20, 30), mode="fill")
        got = cutout.Cutout(ints, None, blc=(-3, 2), trc=(20, 30), mode="fill")
    assert np.array_equal(got.data, want["data"])
    # find_peak on a big-endian surface
    surf = rng.standard_normal((20, 20))
    surf[7, 9] += 10
    assert centroid.find_peak(surf.astype(">f8")) == centroid.find_peak(surf)


def test_ragged_batch_with_an_oversize_pair():
    """One extended source (> 256 px) must not abort the batch."""
    from subpixal_b200 import batch, _lib as L
    shapes = [(32, 32), (10, 300), (40, 24), (2, 30), (64, 64)]
    stacks = [orc.random_pairs(1, ny, nx, seed=50 + k)[0][0] for k, (ny, nx) in enumerate(shapes)]
    flat = np.concatenate([s.ravel() for s in stacks])
    off = np.concatenate([[0], np.cumsum([s.size for s in stacks])])[:-1]
    sh, pk, st, icc, icc_off = batch.cc_fit_ragged(torch.from_numpy(flat).cuda(), off,
                                                   np.array(shapes), "NCC", want_icc=True)
    sh, pk, st, icc = sh.cpu().numpy(), pk.cpu().numpy(), st.cpu().numpy(), icc.cpu().numpy()
    for k, (ny, nx) in enumerate(shapes):
        surf = icc[int(icc_off[k]):int(icc_off[k + 1])]
        if min(ny, nx) < 3 or max(ny, nx) > 256:
            assert np.all(np.isnan(sh[k])) and pk[k] == -1 and st[k] == L.S_UNSUPPORTED_SIZE
            assert np.all(np.isnan(surf))
            continue
        with warnings.catch_warnings(), np.errstate(all="ignore"):
            warnings.simplefilter("ignore")
            dx, dy, oicc, _ = orc.find_displacement(*stacks[k], cc_type="NCC", full_output=True)
        assert abs(sh[k, 0] - dx) < 1e-4 and abs(sh[k, 1] - dy) < 1e-4
        assert int(np.argmax(oicc)) == pk[k]
    with pytest.raises(ValueError):                          # negative offsets are rejected
        batch.cc_fit_ragged(torch.from_numpy(flat).cuda(), off - 5, np.array(shapes), "NCC")


def test_frame_displacements_dq_and_oversize():
    """drz_dq / drz_dqbitmask zero the drizzled pixels with bad DQ bits (create_cutouts
    -> dzct.mask -> align.py:661): identical to a drizzled frame with those pixels zeroed."""
    from subpixal_b200 import align, _lib as L
    rng = np.random.default_rng(8)
    H, W, n = 300, 360, 12
    img = rng.standard_normal((H, W)).astype(np.float32)
    drz = (img + 0.1 * rng.standard_normal((H, W))).astype(np.float32)
    dq = rng.integers(0, 16, (H, W)).astype(np.int16)
    dq[rng.uniform(size=(H, W)) < 0.9] = 0
    x1 = rng.integers(10, W - 80, n); y1 = rng.integers(10, H - 80, n)
    w = rng.integers(20, 60, n); h = rng.integers(20, 60, n)
    ib = np.stack([x1, y1, x1 + w - 1, y1 + h - 1], 1)
    db = ib + np.array([-3, -3, 3, 3])
    fmap = lambda x, y: (x + 0.2, y - 0.1)
    t = lambda a: torch.from_numpy(a).cuda()
    for bits in (0, 4, "~1"):
        from subpixal_b200.cutout import interpret_bit_flags
        good = interpret_bit_flags(bits)
        zeroed = drz.copy()
        zeroed[(dq & ~np.int16(good)) != 0] = 0
        a = align.frame_displacements(t(img), t(drz), ib, db, fmap, drz_dq=t(dq), drz_dqbitmask=bits)
        b = align.frame_displacements(t(img), t(zeroed), ib, db, fmap)
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and torch.equal(a[2], b[2])
    c = align.frame_displacements(t(img), t(drz), ib, db, fmap, drz_dq=t(dq), drz_dqbitmask=None)
    d = align.frame_displacements(t(img), t(drz), ib, db, fmap)
    assert torch.equal(c[0], d[0])
    assert not torch.equal(a[0], d[0])
    # an oversize source in the table: flagged, the others unchanged
    ib2 = np.concatenate([ib[:5], [[0, 0, 299, 40]], ib[5:]])
    db2 = np.concatenate([db[:5], [[0, 0, 305, 45]], db[5:]])
    e = align.frame_displacements(t(img), t(drz), ib2, db2, fmap, full_output=True)
    f = align.frame_displacements(t(img), t(drz), ib, db, fmap, full_output=True)
    keep = [0, 1, 2, 3, 4, 6, 7, 8, 9, 10, 11, 12]
    assert torch.equal(e[0][keep], f[0]) and torch.equal(e[1][keep], f[1])
    assert bool(torch.isnan(e[0][5]).all()) and int(e[1][5]) == -1 and int(e[2][5]) == L.S_UNSUPPORTED_SIZE
    e2 = align.frame_displacements(t(img), t(drz), ib2, db2, fmap)          # no surfaces
    assert torch.equal(e2[0][keep], d[0]) and bool(torch.isnan(e2[0][5]).all())
    assert torch.equal(e[3][:int(e[4][5])], f[3][:int(f[4][5])])
    assert bool(torch.isnan(e[3][int(e[4][5]):int(e[4][6])]).all())
