"""CPU-only tests: host-side logic of the package, the C-ABI library's symbol
table, and that compute entry points fail loudly without a CUDA device."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT
from oracle import subpixal_oracle as orc

HEADER = os.path.join(ROOT, "include", "subpixal_b200.h")


def test_library_exports_every_declared_symbol():
    import subpixal_b200._lib as L
    lib = L.load()
    text = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    declared = set(re.findall(r"\b(spx_[a-z_0-9]+)\s*\(", text))
    assert declared == set(L.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.spx_abi_version() == L.ABI_VERSION == int(
        re.search(r"#define SPX_ABI_VERSION (\d+)", text).group(1))
    for const in ("E_BADARG", "E_UNSUPPORTED", "E_WORKSPACE", "E_DTYPE", "S_PEAK_EDGE",
                  "S_NO_MAXIMUM", "S_FIT_OUTSIDE", "S_BOX_MISMATCH", "MAX_CUTOUT"):
        m = re.search(r"#define SPX_%s\s+\(?(-?(?:0x)?[0-9a-fA-F]+)\)?" % const, text)
        assert int(m.group(1), 0) == getattr(L, const), const
    assert b"workspace" in lib.spx_error_string(L.E_WORKSPACE)
    assert lib.spx_error_string(0) == b"ok"


def test_workspace_query_and_argument_checks_without_gpu():
    import subpixal_b200._lib as L
    lib = L.load()
    assert lib.spx_cc_workspace_bytes(0, 64, 64, 0) == 0
    assert lib.spx_cc_workspace_bytes(10, 300, 64, 0) == 0
    small, big = (lib.spx_cc_workspace_bytes(n, 64, 64, 0) for n in (1, 100000))
    assert 0 < small < big
    assert lib.spx_cc_workspace_bytes(100000, 64, 64, 1) == 0   # ICC goes to the user buffer
    assert lib.spx_cc_workspace_bytes(4, 128, 128, 1) == 0      # 2-CTA cluster: W in shared memory
    assert lib.spx_cc_workspace_bytes(4, 256, 256, 1) > 0       # W/R slab for the 256 px class
    z = C.c_void_p(0)
    assert lib.spx_cc_fit_stack(z, z, z, z, z, None, 1, 8, 8, 0, z, z, z, z, z, 0, z) == L.E_BADARG
    assert lib.spx_gather_cutouts(None, z) == L.E_BADARG
    assert lib.spx_find_peak(z, 0, 1, 4, 4, z, z, 0, 0, 5, 5, z, z, z, z) == L.E_BADARG


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only behaviour")
def test_compute_fails_loudly_without_cuda():
    from subpixal_b200 import cc, centroid, cutout
    ref, ims = orc.known_shift_case(16)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        cc.find_displacement(ref, *ims)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        centroid.find_peak(ref)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        cutout.Cutout(ref, None, blc=(0, 0), trc=(3, 3))


def test_argument_errors_match_reference():
    from subpixal_b200 import cc, centroid, cutout
    ref, ims = orc.known_shift_case(16)
    with pytest.raises(ValueError, match="All cutouts must have same shape"):
        cc.find_displacement(ref, ims[0][:-1], *ims[1:])
    with pytest.raises(ValueError, match="Both 'xmax' and 'ymax'"):
        centroid.find_peak(ref, xmax=1.0)
    with pytest.raises(TypeError):
        centroid.find_peak(ref, peak_fit_box=(1, 2, 3))
    with pytest.raises(ValueError):
        centroid.find_peak(ref, peak_fit_box=0)
    with pytest.raises(ValueError, match="'trc' cannot be None"):
        cutout.Cutout(None, None)
    with pytest.raises(ValueError, match="mode"):
        cutout.Cutout(ref, None, mode="loose")
    with pytest.raises(cutout.NoOverlapError):
        cutout.Cutout(ref, None, blc=(16, 0), trc=(20, 4), mode="fill")
    with pytest.raises(cutout.PartialOverlapError):
        cutout.Cutout(ref, None, blc=(-1, 0), trc=(4, 4))
    with pytest.raises(ValueError, match="Ill-formed"):
        cutout.Cutout(ref, None, blc=(5, 5), trc=(4, 8), mode="fill")
    assert issubclass(cutout.NoOverlapError, ValueError)
    assert issubclass(cutout.PartialOverlapError, ValueError)


def test_empty_cutout_and_properties():
    from subpixal_b200 import cutout
    c = cutout.Cutout(None, None, blc=(2, 3), trc=(9, 7), src_id=5, exptime=10.0)
    assert c.data.shape == (5, 8) and c.data.dtype == np.float32
    assert np.all(np.isnan(c.data)) and np.all(c.mask)
    assert (c.width, c.height, c.naxis, c.naxis1, c.naxis2) == (8, 5, [8, 5], 8, 5)
    assert c.src_pos == (5.5, 5.0) and c.cutout_src_pos == (3.5, 2.0)
    c.cutout_src_pos = (1.0, 1.5)
    assert c.src_pos == (3.0, 4.5)
    assert np.array_equal(c.get_bbox('blc'), [[-.5, -.5], [-.5, 4.5], [7.5, 4.5], [7.5, -.5]])
    assert np.array_equal(c.get_bbox(), [[1.5, 2.5], [1.5, 7.5], [9.5, 7.5], [9.5, 2.5]])
    c.dx -= 0.5
    assert c.dx == -0.5 and c.dy == 0
    c.data = np.zeros((5, 8))
    with pytest.raises(ValueError):
        c.data = np.zeros((4, 8))
    with pytest.raises(ValueError):
        c.exptime = 0
    with pytest.raises(ValueError):
        c.data_units = "flux"
    with pytest.raises(ValueError):
        c.src_weight = -1.0
    with pytest.raises(ValueError):
        c.pscale

    class W:
        def all_pix2world(self, x, y, o):
            return x * 2.0, y * 2.0

        def all_world2pix(self, r, d, o, **kw):
            return r / 2.0, d / 2.0
    c2 = cutout.Cutout(None, W(), blc=(2, 3), trc=(9, 7))
    w = c2.pix2world([[1.0, 1.0]])
    assert np.allclose(w, [[6.0, 8.0]])
    assert np.allclose(c2.world2pix(w), [[1.0, 1.0]])
    c2.dx = -0.5
    assert np.allclose(c2.pix2world([0.0], [0.0]), [[5.0], [6.0]])


def test_slices_and_bit_flags():
    from subpixal_b200 import cutout
    for blc, trc in [((-10, -5), (53, 58)), ((2, 1), (9, 12)), ((90, 95), (120, 99))]:
        ex, ins = cutout._overlap_slices(100, 100, blc, trc)
        o = orc.crop(np.zeros((100, 100), np.float32), blc, trc, mode="fill", fillval=0)
        assert ex == o["extraction_slice"] and ins == o["insertion_slice"]
    f = cutout.interpret_bit_flags
    assert f(None) is None and f("None") is None and f(6) == 6
    assert f("4,8") == 12 and f("4+8") == 12 and f("~4+8") == ~12 and f("~0") == -1
    assert f(-13) == ~12


def test_py2round_and_cc_type_codes():
    from subpixal_b200 import batch, utils
    assert utils.py2round(0.5) == 1.0 and utils.py2round(-0.5) == -1.0
    assert list(utils.py2round(np.array([0.5, -1.5, 1.49]))) == [1.0, -2.0, 1.0]
    assert batch.cc_type_code("zncc") == 2 and batch.cc_type_code("NCC") == 1
    assert batch.cc_type_code("bogus") == 0 and batch.cc_type_code("cc") == 0


def test_shard_range_covers_everything():
    from subpixal_b200 import dist as d
    for n in (0, 1, 7, 8, 9, 10000, 100003):
        for world in (1, 2, 4, 8):
            got, per = [], None
            for r in range(world):
                lo, hi, per = d.shard_range(n, r, world)
                assert 0 <= hi - lo <= per
                got += list(range(lo, hi))
            assert got == list(range(n))
            assert per * world >= n


def test_pipeline_table_packing_is_aligned_and_ordered():
    """align._Tables: every per-source table of frame_displacements lands 16-byte aligned
    in ONE staging buffer, in insertion order (the device pointers are base + offset)."""
    from subpixal_b200.align import _Tables
    t = _Tables()
    a = t.add(np.arange(7), np.int32)             # 28 bytes -> next table starts at 32
    b = t.add(np.arange(3), np.float64)
    c = t.add(np.zeros((5, 20)), np.float64)
    d = t.add(np.arange(2), np.int64)
    assert (a, b, c) == (0, 32, 64) and d == 64 + 800 and t.size == d + 16
    assert all(off % 16 == 0 for off, _ in t.parts)
    assert [arr.dtype for _, arr in t.parts] == [np.int32, np.float64, np.float64, np.int64]


def test_size_classes_of_the_python_side_match_the_library():
    """batch._SIZE_CLASSES lists the class boundaries of pick_cfg() in csrc/spx_cc.cu; the
    boundaries of the classes that keep W / R in a global slab show in the workspace size."""
    import subpixal_b200._lib as L
    from subpixal_b200 import batch
    lib = L.load()
    ws = lambda n: lib.spx_cc_workspace_bytes(1, n, n, 0)
    assert batch._SIZE_CLASSES[-1] == L.MAX_CUTOUT and 128 in batch._SIZE_CLASSES and 170 in batch._SIZE_CLASSES
    assert ws(129) > ws(128) + 500000          # 129 px: first class with a W / R slab (P = 256)
    assert ws(171) > ws(170) + 500000          # 171 px: P = 384
    assert ws(170) - ws(129) < 300000          # inside a class only the surface scratch grows
    assert lib.spx_cc_workspace_bytes(1, L.MAX_CUTOUT + 1, 8, 0) == 0


def test_native_array_and_device_frame_conversions():
    """FITS arrays are big-endian; integer frames are accepted like the
    reference does (cutout.py:737-752) - host-side conversions only."""
    from subpixal_b200.utils import native_array
    from subpixal_b200 import cutout
    be = np.arange(12, dtype=">f4").reshape(3, 4)
    nat = native_array(be)
    assert nat.dtype.isnative and nat.flags.c_contiguous and np.array_equal(nat, be)
    torch.from_numpy(nat)                                   # would raise on '>f4'
    host, restore = cutout._device_frame(be)
    assert host.dtype == np.float32 and host.dtype.isnative and restore == be.dtype
    host, restore = cutout._device_frame(np.arange(6, dtype=">i2").reshape(2, 3))
    assert host.dtype == np.float64 and restore == np.dtype(">i2")
    assert np.array_equal(host, [[0, 1, 2], [3, 4, 5]])
    host, restore = cutout._device_frame(np.ones((2, 2), dtype=np.float32))
    assert restore is None
    with pytest.raises(TypeError):
        cutout._device_frame(np.full((2, 2), 1 << 60, dtype=np.int64))
    with pytest.raises(TypeError):
        cutout._device_frame(np.ones((2, 2), dtype=np.complex64))
    view, dt = cutout._dq_device_view(np.arange(4, dtype=">u2").reshape(2, 2))
    assert view.dtype == np.int16 and view.dtype.isnative and dt.isnative
    torch.from_numpy(view)


def test_scratch_cache_release():
    from subpixal_b200 import batch
    dev = torch.device("cpu")

    class FakeStream:
        cuda_stream = 0
    orig = torch.cuda.current_stream
    torch.cuda.current_stream = lambda d=None: FakeStream()
    try:
        big = batch.scratch_f32("t", 64 << 20, dev)
        assert batch.scratch_f32("t", 60 << 20, dev) is big          # reused
        small = batch.scratch_f32("t", 1 << 20, dev)                 # far too large: dropped
        assert small is not big and small.numel() == 1 << 20
        batch.clear_scratch()
        assert not batch._scratch_bufs and not batch._workspaces
    finally:
        torch.cuda.current_stream = orig


def test_public_surface_matches_the_reference():
    """Drop-in guard: every public attribute of the reference's Cutout, the builders'
    signatures and the signatures of find_displacement / find_peak / parse_file_name are
    the reference's own (checked live when the reference sources are present)."""
    import inspect
    from oracle import refload
    if not refload.available():
        pytest.skip("reference tree not present")
    import subpixal_b200.cutout as mc
    import subpixal_b200.cc as mcc
    import subpixal_b200.centroid as mce
    import subpixal_b200.ingest as mi
    rc, rcc, rce, ru = (refload.load(n) for n in ("cutout", "cc", "centroid", "utils"))
    public = lambda o: {n for n in dir(o) if not n.startswith("_")}      # noqa: E731
    assert public(rc.Cutout) == public(mc.Cutout)
    assert all(hasattr(mc, n) for n in rc.__all__)
    for name in ("create_primary_cutouts", "create_cutouts", "create_input_image_cutouts",
                 "drz_from_input_cutouts"):
        assert str(inspect.signature(getattr(rc, name))) == str(inspect.signature(getattr(mc, name)))
    assert str(inspect.signature(rc.Cutout.__init__)) == str(inspect.signature(mc.Cutout.__init__))
    assert str(inspect.signature(rcc.find_displacement)) == str(inspect.signature(mcc.find_displacement))
    assert str(inspect.signature(rce.find_peak)) == str(inspect.signature(mce.find_peak))
    assert str(inspect.signature(ru.parse_file_name)) == str(inspect.signature(mi.parse_file_name))
    assert issubclass(mc.NoOverlapError, ValueError) and issubclass(mc.PartialOverlapError, ValueError)


class _FakeCutout(object):
    """What blot / align need of a Cutout for the host-side WCS arithmetic (no device)."""
    DEFAULT_ACCURACY, DEFAULT_MAXITER, DEFAULT_QUIET = 1e-5, 50, True

    def __init__(self, wcs, blc, naxis, dx=0.0, dy=0.0):
        self.wcs, self.blc, self.naxis, self.dx, self.dy = wcs, blc, naxis, dx, dy

    def pix2world(self, x, y, origin=0):
        x = np.asarray(x, float) + (self.blc[0] - self.dx)
        y = np.asarray(y, float) + (self.blc[1] - self.dy)
        return list(self.wcs.all_pix2world(x, y, origin))

    def world2pix(self, ra, dec, origin=0):
        x, y = self.wcs.all_world2pix(ra, dec, origin, accuracy=1e-5, maxiter=50, quiet=True)
        return [x - (self.blc[0] - self.dx), y - (self.blc[1] - self.dy)]


def test_batched_pixel_maps_equal_the_per_pair_maps():
    """blot_phases fits the 9-point maps of all pairs that share their two WCS objects with
    ONE call of each WCS: bit-identical to fit_pixel_map(BlotWCSMap(target, source)) per pair
    (blot.py:21-76, 119-120), including half-pixel shifted targets and 1-pixel axes."""
    from wcs_stubs import AffineWCS, QuadWCS
    from subpixal_b200 import blot
    rng = np.random.default_rng(0)
    tw = [QuadWCS([[1, 0.01], [-0.01, 1]], [5.0, 7.0], [100.0, 100.0]),
          AffineWCS([[0.99, 0.0], [0.0, 1.01]], [1.0, 2.0])]
    sw = AffineWCS(np.eye(2) * 0.9, [3.0, 2.0])
    tcs = [_FakeCutout(tw[k % 2], (int(rng.integers(0, 200)), int(rng.integers(0, 200))),
                       (int(rng.integers(1, 70)), int(rng.integers(1, 70))),
                       dx=float(rng.choice([0, 0.5])), dy=float(rng.choice([0, 0.5])))
           for k in range(40)]
    scs = [_FakeCutout(sw, (int(rng.integers(0, 200)), int(rng.integers(0, 200))), (40, 40))
           for _ in range(40)]
    got = blot._cutout_pixel_maps(scs, tcs)
    want = np.stack([blot.fit_pixel_map(blot.BlotWCSMap(t, s), int(t.naxis[0]), int(t.naxis[1]))
                     for s, t in zip(scs, tcs)])
    assert np.array_equal(got, want)


def test_cutout_deepcopy_copies_arrays_and_honours_the_memo():
    """Cutout.__deepcopy__: independent arrays and lists, a deep copy of the WCS unless the
    caller shares it through ``memo`` (find_linear_fit does, for its blot copies)."""
    import copy
    from subpixal_b200.cutout import Cutout

    class Wcs(object):
        def __init__(self):
            self.v = [1, 2]
    w = Wcs()
    c = Cutout.__new__(Cutout)
    c.__dict__.update(dict(
        _onx=10, _ony=12, _blc=(1, 2), _trc=(5, 6), _src_pos=(3.0, 4.0), _src_weight=None,
        _data=np.arange(25.0).reshape(5, 5), _mask=np.zeros((5, 5), bool), _dq=None, _weight=None,
        _src_id=3, _dx=0, _dy=-0.5, _wcs=w, _naxis=[5, 5], naxis1=5, naxis2=5,
        _eslice=(slice(2, 7), slice(1, 6)), _islice=(slice(0, 5), slice(0, 5)), _exptime=1,
        _data_units='rate'))
    d = copy.deepcopy(c)
    assert d is not c and d._data is not c._data and np.array_equal(d._data, c._data)
    assert d._wcs is not w and d._wcs.v == [1, 2] and d._naxis is not c._naxis and d._naxis == [5, 5]
    assert d._eslice == c._eslice and d._dy == -0.5 and d._data_units == 'rate'
    d._data[0, 0] = 99
    assert c._data[0, 0] == 0
    assert copy.deepcopy(c, {id(w): w})._wcs is w


def test_every_python_file_compiles():
    """bench.py, __graft_entry__.py, the package, the oracle and every developer tool parse."""
    import glob
    files = [os.path.join(ROOT, "bench.py"), os.path.join(ROOT, "__graft_entry__.py")]
    for sub in ("subpixal_b200", "oracle", "tools", "tests", os.path.join("tests", "golden")):
        files += glob.glob(os.path.join(ROOT, sub, "*.py"))
    assert len(files) > 40
    for f in files:
        with open(f) as fh:
            compile(fh.read(), f, "exec")
