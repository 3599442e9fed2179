"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED
reference (/root/reference/subpixal/{cc,centroid,cutout}.py) on deterministic
synthetic inputs.  Run in the build container only:

    python tests/golden/make_golden.py

The reference has no tests or fixtures of its own, so
these files are the parity pin: reference source + numpy 2.3.5 / scipy 1.18.1
of this image.  Fixtures are small .npz files (inputs included, so nothing has
to be regenerated bit-exactly on the GPU box).
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import refload                      # noqa: E402
from oracle import subpixal_oracle as orc       # noqa: E402

CC_TYPES = ("CC", "NCC", "ZNCC")


def run_fd(cc, stack, cc_type):
    with warnings.catch_warnings(), np.errstate(all="ignore"):
        warnings.simplefilter("ignore")
        dx, dy, icc, _ = cc.find_displacement(stack[0], stack[1], stack[2],
                                              stack[3], stack[4],
                                              cc_type=cc_type,
                                              full_output=True)
    with np.errstate(all="ignore"):
        am = int(np.argmax(icc.astype(np.float64)))
    return float(dx), float(dy), am, icc


def known_shift(cc):
    rows = []
    for n in (32, 64, 128, 256):
        ref, ims = orc.known_shift_case(n)
        st = np.stack([ref] + ims)
        for t, name in enumerate(CC_TYPES):
            dx, dy, am, _ = run_fd(cc, st, name)
            rows.append((n, t, dx, dy, am))
    np.savez(os.path.join(HERE, "known_shift.npz"),
             table=np.array(rows, dtype=np.float64),
             cc_types=np.array(CC_TYPES))


def random_cases(cc):
    specs = {                      # name: (n_pairs, ny, nx, seed, masked)
        "rand32": (8, 32, 32, 20241, False),
        "rand64": (3, 64, 64, 20242, False),
        "rand33": (4, 33, 33, 20243, False),
        "rect20x37": (4, 20, 37, 20244, False),
        "mask48": (3, 48, 48, 20245, True),
        "rand128": (1, 128, 128, 20246, False),
    }
    for name, (n, ny, nx, seed, masked) in specs.items():
        stacks, _ = orc.random_pairs(n, ny, nx, seed=seed, masked=masked)
        res = {}
        for t in CC_TYPES:
            out = [run_fd(cc, s, t) for s in stacks]
            res["dxdy_" + t] = np.array([[o[0], o[1]] for o in out])
            res["argmax_" + t] = np.array([o[2] for o in out], dtype=np.int64)
            if ny <= 64 or t == "NCC":        # keep the big fixture small
                res["icc_" + t] = np.stack([o[3] for o in out])
        np.savez_compressed(os.path.join(HERE, name + ".npz"),
                            stacks=stacks, **res)


def degenerate(cc):
    """NaN / all-zero / unknown cc_type behaviour."""
    n = 32
    base, _ = orc.random_pairs(1, n, n, seed=20250)
    base = base[0]
    cases, names = [], []

    def add(name, st, t):
        dx, dy, am, icc = run_fd(cc, st, t)
        nanfrac = float(np.mean(np.isnan(icc)))
        cases.append((dx, dy, am, nanfrac))
        names.append(name + ":" + t)
        return st

    stacks = []
    for slot in range(5):
        st = base.copy()
        st[slot, 5, 7] = np.nan
        for t in ("CC", "NCC"):
            add("nan_slot%d" % slot, st, t)
        stacks.append(st)
    st = base.copy()
    st[1:] = 0.0
    for t in CC_TYPES:
        add("zero_phases", st, t)
    stacks.append(st)
    st = base.copy()
    st[3, 4, 4] = np.inf
    for t in ("CC", "NCC"):
        add("inf_slot3", st, t)
    stacks.append(st)
    add("bogus_type", base, "bogus")
    add("lower_case", base, "zncc")
    stacks.append(base)
    np.savez_compressed(os.path.join(HERE, "degenerate.npz"),
                        stacks=np.stack(stacks),
                        results=np.array(cases), names=np.array(names))


def peak_cases(ce):
    rng = np.random.default_rng(20260)
    imgs, outs, kw = [], [], []

    def g16(x0, y0):
        return orc.gaussian_star(16, x0, y0, 2.0, amp=1.0).astype(np.float64)

    for (x0, y0) in [(0, 5), (1, 5), (14, 5), (15, 5), (14.4, 5), (15, 15),
                     (7.3, 8.6), (5, 0), (5, 1), (5, 14), (5, 15), (1, 1),
                     (2.2, 13.1)]:
        imgs.append(g16(x0, y0))
    imgs.append(np.ones((16, 16)))
    imgs.append(np.zeros((16, 16)))
    two = np.zeros((16, 16))
    two[4, 9] = two[9, 4] = 1.0
    imgs.append(two + 0.01 * g16(8, 8))
    nanimg = g16(8, 8)
    nanimg[3, 11] = np.nan
    imgs.append(nanimg)
    imgs.append(-g16(8, 8))                     # no maximum: bowl
    for _ in range(6):
        imgs.append(g16(*rng.uniform(2, 13, 2)) +
                    0.02 * rng.standard_normal((16, 16)))
    for im in imgs:
        with np.errstate(all="ignore"):
            outs.append(ce.find_peak(im, peak_fit_box=5,
                                     peak_search_box="all"))
    # keyword variants on a fixed noisy image
    im = g16(9.3, 6.7) + 0.02 * rng.standard_normal((16, 16))
    msk = rng.uniform(size=(16, 16)) > 0.15
    variants = [
        dict(),
        dict(peak_fit_box=3),
        dict(peak_fit_box=(7, 5)),
        dict(peak_fit_box=2),
        dict(xmax=9.0, ymax=7.0),
        dict(xmax=8.5, ymax=6.5, peak_search_box=5),
        dict(xmax=3.0, ymax=3.0, peak_search_box=(3, 3)),
        dict(xmax=3.0, ymax=3.0, peak_search_box="fitbox"),
        dict(xmax=3.0, ymax=3.0, peak_search_box="all"),
        dict(xmax=3.0, ymax=3.0, peak_search_box="off"),
        dict(mask=msk),
        dict(mask=msk, xmax=9.0, ymax=7.0, peak_search_box=7),
        dict(xmax=-0.5, ymax=2.5),
    ]
    vouts = []
    for v in variants:
        vouts.append(ce.find_peak(im, **v))
        kw.append(repr({k: (val if k != "mask" else "MASK")
                        for k, val in v.items()}))
    np.savez_compressed(os.path.join(HERE, "find_peak.npz"),
                        images=np.stack(imgs), peaks=np.array(outs),
                        var_image=im, var_mask=msk, var_kwargs=np.array(kw),
                        var_peaks=np.array(vouts, dtype=np.float64))


def cutout_cases(ct):
    rng = np.random.default_rng(20270)
    ny, nx = 40, 50
    frame = rng.standard_normal((ny, nx)).astype(np.float32)
    frame[3, 4] = np.nan
    frame[20, 30] = np.inf
    frame[39, 49] = -np.inf
    dq = rng.integers(0, 16, size=(ny, nx)).astype(np.int16)
    wht = rng.uniform(size=(ny, nx)).astype(np.float32)
    boxes = [(0, 0, 49, 39), (-10, -5, 53, 58), (2, 1, 9, 12), (25, 15, 60, 30),
             (-3, 30, 4, 45), (48, 38, 49, 39), (10, 10, 10, 10)]
    out = {"frame": frame, "dq": dq, "wht": wht,
           "boxes": np.array(boxes, dtype=np.int64)}
    for k, (x1, y1, x2, y2) in enumerate(boxes):
        c = ct.Cutout(frame, None, blc=(x1, y1), trc=(x2, y2), dq=dq,
                      weight=wht, mode="fill", fillval=0)
        out["data%d" % k] = c.data
        out["mask%d" % k] = c.mask
        out["dq%d" % k] = c.dq
        out["wht%d" % k] = c.weight
        es, is_ = c.extraction_slice, c.insertion_slice
        out["slices%d" % k] = np.array(
            [es[0].start, es[0].stop, es[1].start, es[1].stop,
             is_[0].start, is_[0].stop, is_[1].start, is_[1].stop])
    np.savez_compressed(os.path.join(HERE, "cutout.npz"), **out)


def builder_cases(ct):
    """The reference's OWN list builders (cutout.py:47-198, 201-306, 309-438, 441-587)
    on the scenes of tests/wcs_stubs.py: boxes, masks, source positions, list order
    and the dropped sources are frozen for the product's builders to match exactly."""
    sys.path.insert(0, os.path.dirname(HERE))
    import wcs_stubs as ws
    out = {}
    for seed in (20280, 20281):
        sc = ws.builder_scene(seed)
        for name, pkw, ckw in ws.BUILDER_CASES:
            if not ws.case_in_fixture(seed, name):
                continue
            prim, flt_cts, drz_cts = ws.run_builders(ct, sc, pkw, ckw)
            tag = "s%d_%s" % (seed, name)
            ws.pack_cutouts(tag + "_prim", prim, out)
            ws.pack_cutouts(tag + "_flt", flt_cts, out)
            ws.pack_cutouts(tag + "_drz", drz_cts, out)
            out[tag + "_pscale"] = np.array([flt_cts[0].pscale, drz_cts[0].pscale])
    np.savez_compressed(os.path.join(HERE, "builders.npz"), **out)


def main():
    if not refload.available():
        raise SystemExit("reference tree not mounted; fixtures unchanged")
    cc, ce, ct = (refload.load(m) for m in ("cc", "centroid", "cutout"))
    known_shift(cc)
    random_cases(cc)
    degenerate(cc)
    peak_cases(ce)
    cutout_cases(ct)
    builder_cases(ct)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
