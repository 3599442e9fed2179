"""Pin the CPU oracle: (1) against the committed golden fixtures produced by
the unmodified reference (tests/golden/make_golden.py), (2) live against the
reference sources when /root/reference is mounted (build container only)."""
import warnings

import numpy as np
import pytest

from conftest import golden
from oracle import refload
from oracle import subpixal_oracle as orc

CC_TYPES = ("CC", "NCC", "ZNCC")
RANDOM_SETS = ("rand32", "rand64", "rand33", "rect20x37", "mask48", "rand128")


def _fd(stack, t):
    with warnings.catch_warnings(), np.errstate(all="ignore"):
        warnings.simplefilter("ignore")
        return orc.find_displacement(*stack, cc_type=t, full_output=True)


def test_known_shift_table():
    g = golden("known_shift")
    for n, t, dx, dy, am in g["table"]:
        ref, ims = orc.known_shift_case(int(n))
        odx, ody, icc, _ = _fd([ref] + ims, CC_TYPES[int(t)])
        assert abs(odx - dx) < 1e-9 and abs(ody - dy) < 1e-9
        assert int(np.argmax(icc)) == int(am)
    # the survey's frozen N=64 values
    ref, ims = orc.known_shift_case(64)
    dx, dy = orc.find_displacement(ref, *ims, cc_type="CC")
    assert abs(dx - 0.371433842) < 1e-8 and abs(dy + 0.207700258) < 1e-8


@pytest.mark.parametrize("name", RANDOM_SETS)
def test_random_sets_bit_exact(name):
    g = golden(name)
    stacks = g["stacks"]
    for t in CC_TYPES:
        for k, st in enumerate(stacks):
            dx, dy, icc, ccs = _fd(st, t)
            assert (dx, dy) == tuple(g["dxdy_" + t][k])
            assert int(np.argmax(icc)) == int(g["argmax_" + t][k])
            key = "icc_" + t
            if key in g.files:
                assert icc.dtype == np.float32
                assert np.array_equal(icc, g[key][k])


def test_degenerate_cases():
    g = golden("degenerate")
    stacks, names, res = g["stacks"], g["names"], g["results"]
    order = ["nan_slot0", "nan_slot1", "nan_slot2", "nan_slot3", "nan_slot4",
             "zero_phases", "inf_slot3", "bogus_type"]
    for name, (dx, dy, am, nanfrac) in zip(names, res):
        case, t = str(name).split(":")
        st = stacks[order.index(case)] if case != "lower_case" else stacks[-1]
        odx, ody, icc, _ = _fd(st, t)
        assert (odx, ody) == (dx, dy), name
        with np.errstate(all="ignore"):
            assert int(np.argmax(icc.astype(np.float64))) == int(am)
        assert float(np.mean(np.isnan(icc))) == nanfrac
    with pytest.raises(ValueError):
        orc.find_displacement(stacks[0][0], stacks[0][1][:-1], *stacks[0][2:])


def test_find_peak_fixtures():
    g = golden("find_peak")
    for im, pk in zip(g["images"], g["peaks"]):
        with np.errstate(all="ignore"):
            got = orc.find_peak(im, peak_fit_box=5, peak_search_box="all")
        assert np.allclose(got, pk, rtol=0, atol=1e-12, equal_nan=True)
    im, msk = g["var_image"], g["var_mask"]
    for kw, pk in zip(g["var_kwargs"], g["var_peaks"]):
        kwargs = eval(str(kw), {"MASK": None})
        if "mask" in kwargs:
            kwargs["mask"] = msk
        got = orc.find_peak(im, **kwargs)
        assert np.allclose(got, pk, rtol=0, atol=1e-12), kw
    with pytest.raises(ValueError):
        orc.find_peak(im, xmax=1.0)
    with pytest.raises(TypeError):
        orc.find_peak(im, peak_fit_box=(1, 2, 3))
    with pytest.raises(ValueError):
        orc.find_peak(im, peak_fit_box=0)


def test_cutout_fixtures():
    g = golden("cutout")
    frame, dq, wht = g["frame"], g["dq"], g["wht"]
    for k, (x1, y1, x2, y2) in enumerate(g["boxes"]):
        c = orc.crop(frame, (x1, y1), (x2, y2), dq=dq, weight=wht,
                     mode="fill", fillval=0)
        assert np.array_equal(c["data"], g["data%d" % k], equal_nan=True)
        assert np.array_equal(c["mask"], g["mask%d" % k])
        assert np.array_equal(c["dq"], g["dq%d" % k])
        assert np.array_equal(c["weight"], g["wht%d" % k])
        es, is_ = c["extraction_slice"], c["insertion_slice"]
        assert [es[0].start, es[0].stop, es[1].start, es[1].stop,
                is_[0].start, is_[0].stop, is_[1].start,
                is_[1].stop] == list(g["slices%d" % k])
    # blc=(-10,-5), trc=(53,58) -> 910 masked... on a
    # frame that covers it except for the negative part
    big = np.zeros((100, 100), dtype=np.float32)
    c = orc.crop(big, (-10, -5), (53, 58), mode="fill", fillval=0)
    assert int(c["mask"].sum()) == 910
    with pytest.raises(orc.PartialOverlap):
        orc.crop(frame, (-1, 0), (5, 5), mode="strict")
    with pytest.raises(orc.NoOverlap):
        orc.crop(frame, (50, 0), (55, 5), mode="fill")
    with pytest.raises(ValueError):
        orc.crop(frame, (5, 5), (4, 8), mode="fill")


def test_icc_direct_identity():
    """The FFT oracle equals the direct lag-sum definition (not an FFT)."""
    stacks, _ = orc.random_pairs(2, 12, 15, seed=7)
    for st in stacks:
        icc, _ = orc.build_icc(*st.astype(np.float64), cc_type="CC")
        ref = orc.icc_direct(st[0], st[1:])
        assert np.allclose(icc, ref, rtol=1e-10, atol=1e-8)


def test_py2round():
    assert orc.py2round(0.5) == 1.0 and orc.py2round(-0.5) == -1.0
    assert orc.py2round(2.4) == 2.0 and orc.py2round(-2.6) == -3.0
    assert list(orc.py2round(np.array([0.5, -1.5, 1.49]))) == [1.0, -2.0, 1.0]


@pytest.mark.skipif(not refload.available(),
                    reason="reference tree only exists in the build container")
def test_live_against_reference():
    cc, ce, ct = (refload.load(m) for m in ("cc", "centroid", "cutout"))
    stacks, _ = orc.random_pairs(3, 24, 31, seed=99, masked=True)
    for st in stacks:
        for t in CC_TYPES + ("weird",):
            with warnings.catch_warnings(), np.errstate(all="ignore"):
                warnings.simplefilter("ignore")
                a = cc.find_displacement(*st, cc_type=t, full_output=True)
                b = orc.find_displacement(*st, cc_type=t, full_output=True)
            assert a[0] == b[0] and a[1] == b[1]
            assert np.array_equal(a[2], b[2])
            for x, y in zip(a[3], b[3]):
                assert np.array_equal(x, y)
    rng = np.random.default_rng(5)
    for _ in range(20):
        im = rng.standard_normal((9, 11))
        kw = dict(peak_fit_box=int(rng.integers(2, 6)))
        if rng.uniform() < 0.5:
            kw.update(xmax=float(rng.uniform(0, 10)),
                      ymax=float(rng.uniform(0, 8)),
                      peak_search_box=int(rng.integers(1, 6)))
        assert np.allclose(ce.find_peak(im, **kw), orc.find_peak(im, **kw),
                           rtol=0, atol=1e-12)
    frame = rng.standard_normal((30, 20)).astype(np.float32)
    c = ct.Cutout(frame, None, blc=(-2, 3), trc=(7, 33), mode="fill",
                  fillval=0)
    o = orc.crop(frame, (-2, 3), (7, 33), mode="fill", fillval=0)
    assert np.array_equal(c.data, o["data"]) and np.array_equal(c.mask,
                                                                 o["mask"])
    dq = rng.integers(0, 64, size=(4, 4)).astype(np.int32)
    ref_bad = refload._bitfield_to_boolean_mask(dq, 6, good_mask_value=False)
    assert np.array_equal(ref_bad, orc.bitfield_bad_mask(dq, 6))


@pytest.mark.skipif(not refload.available(),
                    reason="/root/reference only exists in the build container")
def test_builder_fixture_is_what_the_reference_produces():
    """tests/golden/builders.npz = outputs of the reference's own create_primary_cutouts /
    create_cutouts (cutout.py:47-587) on the scenes of tests/wcs_stubs.py, re-run live."""
    import wcs_stubs as ws
    ct = refload.load("cutout")
    g = golden("builders")
    seed = ws.BUILDER_SEEDS[1]
    for name, pkw, ckw in ws.BUILDER_CASES:
        if not ws.case_in_fixture(seed, name):
            continue
        out = {}
        prim, flt_cts, drz_cts = ws.run_builders(ct, ws.builder_scene(seed), pkw, ckw)
        tag = "s%d_%s" % (seed, name)
        for suffix, cts in (("_prim", prim), ("_flt", flt_cts), ("_drz", drz_cts)):
            ws.pack_cutouts(tag + suffix, cts, out)
        assert out and all(np.array_equal(v, g[k], equal_nan=True) for k, v in out.items())
