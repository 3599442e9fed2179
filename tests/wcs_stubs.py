"""Stand-ins for astropy objects used by the cutout builders' fixtures: WCS objects
with astropy's calling conventions (``all_pix2world`` / ``all_world2pix`` with either
``(Nx2 array, origin)`` or ``(x, y, origin)``) and a table exposing ``colnames``.
Shared by tests/golden/make_golden.py (which feeds them to the UNMODIFIED reference
builders, cutout.py:47-587) and by the GPU tests of the product builders."""
import numpy as np


class AffineWCS(object):
    """world = A @ pix + b (0-based pixels)."""

    def __init__(self, a, b):
        self.a, self.b = np.asarray(a, float), np.asarray(b, float)
        self.ai = np.linalg.inv(self.a)
        self.pixel_scale_matrix = self.a

    @staticmethod
    def _args(args):
        if len(args) == 2:                       # (Nx2 array, origin)
            arr = np.asarray(args[0], float)
            return arr[:, 0], arr[:, 1], args[1], True
        return (np.atleast_1d(np.asarray(args[0], float)), np.atleast_1d(np.asarray(args[1], float)),
                args[2], False)

    def _fwd(self, p):
        return p @ self.a.T + self.b

    def _inv(self, w):
        return (w - self.b) @ self.ai.T

    def all_pix2world(self, *args, **kw):
        x, y, o, packed = self._args(args)
        w = self._fwd(np.stack([x - o, y - o], 1))
        return w if packed else (w[:, 0], w[:, 1])

    def all_world2pix(self, *args, **kw):
        r, d, o, packed = self._args(args)
        p = self._inv(np.stack([r, d], 1)) + o
        return p if packed else (p[:, 0], p[:, 1])


class QuadWCS(AffineWCS):
    """Affine map plus a small quadratic distortion (a detector-like, non-linear WCS);
    the inverse is a fixed-point iteration converged far below the 1e-5 px the
    builders ask of ``all_world2pix`` (cutout.py:265-266)."""

    def __init__(self, a, b, centre, eps=2.0e-5):
        super().__init__(a, b)
        self.c, self.eps = np.asarray(centre, float), float(eps)

    def _dist(self, p):
        d = p - self.c
        q = np.stack([d[:, 0] * d[:, 1] + 0.5 * d[:, 1] ** 2, d[:, 0] ** 2 - 0.3 * d[:, 0] * d[:, 1]], 1)
        return p + self.eps * q

    def _fwd(self, p):
        return self._dist(p) @ self.a.T + self.b

    def _inv(self, w):
        lin = (w - self.b) @ self.ai.T
        p = lin.copy()
        for _ in range(60):
            p = lin - (self._dist(p) - p)
        return p


class Table(object):
    """Minimal astropy.table.Table look-alike: ``colnames`` + column access."""

    def __init__(self, **cols):
        self._cols = {k: np.asarray(v) for k, v in cols.items()}
        self.colnames = list(self._cols)
        self.meta = {}

    def __getitem__(self, name):
        return self._cols[name]

    def keys(self):
        return self._cols.keys()


def builder_scene(seed, ny=170, nx=190, nsrc=16, fny=160, fnx=180):
    """Deterministic inputs of the builder fixtures: a drizzled image with a
    segmentation image, DQ and weight planes, a catalogue (two ids without usable
    segments), a distorted 'flt' image with its own DQ plane, and the two WCSs."""
    rng = np.random.default_rng(seed)
    drz = rng.standard_normal((ny, nx)).astype(np.float32)
    seg = np.zeros((ny, nx), dtype=np.int32)
    ids, xs, ys = [], [], []
    for k in range(nsrc):
        sid = 3 * k + 2
        w, h = rng.integers(3, 16, 2)
        x1, y1 = rng.integers(4, nx - w - 4), rng.integers(4, ny - h - 4)   # pad <= 3 stays inside
        blob = rng.uniform(size=(h, w)) < 0.8
        seg[y1:y1 + h, x1:x1 + w][blob] = sid
        ids.append(sid); xs.append(x1 + w / 2.0 + rng.uniform(-0.4, 0.4))
        ys.append(y1 + h / 2.0 + rng.uniform(-0.4, 0.4))
    seg[0, 5:9] = 999                              # touches the edge -> skipped
    seg[60:64, nx - 1] = 998
    ids += [999, 998, 12345]                       # 12345 has no segment
    xs += [6.0, nx - 1.0, 1.0]; ys += [0.0, 61.0, 1.0]
    drz[10, 10] = np.nan
    drz[seg == 5][:1] = np.inf
    dq = rng.integers(0, 8, (ny, nx)).astype(np.uint16)
    dq[rng.uniform(size=(ny, nx)) < 0.7] = 0
    wht = rng.uniform(0.5, 2.0, (ny, nx)).astype(np.float32)
    flt = rng.standard_normal((fny, fnx)).astype(np.float32)
    flt[40, 41] = np.nan
    fdq = rng.integers(0, 4, (fny, fnx)).astype(np.int16)
    fdq[rng.uniform(size=(fny, fnx)) < 0.8] = 0
    order = rng.permutation(len(ids))              # the catalogue is not sorted by id
    cat = dict(id=np.array(ids)[order], x=np.array(xs)[order], y=np.array(ys)[order],
               weight=np.linspace(1.0, 2.0, len(ids))[order])
    th = 0.05
    drz_wcs = AffineWCS(np.eye(2) * 0.9, [100.0, 50.0])
    flt_wcs = QuadWCS([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]], [104.0, 47.0],
                      centre=[fnx / 2.0, fny / 2.0])
    return dict(drz=drz, seg=seg, dq=dq, wht=wht, flt=flt, fdq=fdq, cat=cat,
                drz_wcs=drz_wcs, flt_wcs=flt_wcs)


# (name, kwargs of create_primary_cutouts, kwargs of create_cutouts)
BUILDER_CASES = (
    ("plain", dict(pad=1), dict(pad=2)),
    ("dq0", dict(pad=1, use_dq=True, dqbitmask=0), dict(pad=2, use_dq=True, drz_dqbitmask=0,
                                                         flt_dqbitmask=0)),
    ("dq6_pad", dict(pad=2.5, use_dq=True, dqbitmask=6), dict(pad=1.5, use_dq=True, drz_dqbitmask=6,
                                                               flt_dqbitmask=1, combine_seg_mask=False)),
    ("dqnone", dict(pad=0, use_dq=True, dqbitmask=None), dict(pad=3, use_dq=True, drz_dqbitmask=None,
                                                               flt_dqbitmask=None)),
)


BUILDER_SEEDS = (20280, 20281)


def case_in_fixture(seed, name):
    """Keeps tests/golden/builders.npz small: every case on the first scene, two on the second."""
    return seed == BUILDER_SEEDS[0] or name in ("dq0", "dq6_pad")


def run_builders(mod, scene, pkw, ckw):
    """Call ``mod.create_primary_cutouts`` and ``mod.create_cutouts`` (``mod`` = the
    reference's cutout module or the product's) on a scene; returns the three lists."""
    pkw, ckw = dict(pkw), dict(ckw)
    cat = Table(**scene["cat"])
    use_dq = pkw.pop("use_dq", False)
    prim = mod.create_primary_cutouts(
        cat, scene["seg"], scene["drz"], scene["drz_wcs"],
        imdq=scene["dq"] if use_dq else None, imweight=None if use_dq else scene["wht"],
        data_units="rate", exptime=3.0, **pkw)
    use_dq = ckw.pop("use_dq", False)
    flt_cts, drz_cts = mod.create_cutouts(
        prim, scene["seg"], scene["drz"], scene["drz_wcs"], scene["flt"], scene["flt_wcs"],
        drz_dq=scene["dq"] if use_dq else None, drz_weight=None if use_dq else scene["wht"],
        flt_dq=scene["fdq"] if use_dq else None, drz_exptime=3.0, flt_exptime=2.0, **ckw)
    return prim, flt_cts, drz_cts


def pack_cutouts(prefix, cts, out):
    """Flatten a list of Cutout objects into npz-able arrays (ragged pixels keyed per cutout)."""
    out[prefix + "_n"] = np.array(len(cts))
    out[prefix + "_id"] = np.array([c.src_id for c in cts], dtype=np.int64)
    out[prefix + "_box"] = np.array([list(c.blc) + list(c.trc) for c in cts], dtype=np.int64).reshape(-1, 4)
    out[prefix + "_src_pos"] = np.array([c.src_pos for c in cts], dtype=np.float64).reshape(-1, 2)
    out[prefix + "_cutout_src_pos"] = np.array([c.cutout_src_pos for c in cts],
                                               dtype=np.float64).reshape(-1, 2)
    out[prefix + "_weight"] = np.array([np.nan if c.src_weight is None else c.src_weight for c in cts],
                                       dtype=np.float64)
    out[prefix + "_exptime"] = np.array([c.exptime for c in cts], dtype=np.float64)
    for k, c in enumerate(cts):
        out["%s_data%d" % (prefix, k)] = np.asarray(c.data)
        out["%s_mask%d" % (prefix, k)] = np.asarray(c.mask)
        if c.dq is not None:
            out["%s_dq%d" % (prefix, k)] = np.asarray(c.dq)
        if c.weight is not None:
            out["%s_wht%d" % (prefix, k)] = np.asarray(c.weight)
