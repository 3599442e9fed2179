"""Loader for the UNMODIFIED reference sources (test infrastructure only).

This module is part of ``oracle/`` and therefore test infrastructure: it may be
imported only from ``tests/``, ``tests/golden/make_golden.py`` and similar
checker scripts - never from the product package ``subpixal_b200``.

The reference (``/root/reference/subpixal``) is pure Python but cannot be
imported as a package in this image: ``subpixal/__init__.py:8-17`` pulls a
generated ``version.py`` plus astropy / stwcs / drizzlepac / stsci.tools, none
of which are installed, and ``cc.py:150`` uses ``np.bool`` (removed in
numpy >= 1.24).  The hot-path files themselves (``utils.py``, ``centroid.py``,
``cc.py``, ``cutout.py``) only *use* numpy and scipy, so we execute them by
path with stub modules registered for the missing imports (recipe of
SURVEY.md section 8c).  ``/root/reference`` only exists in the build
container, so ``available()`` must be checked first; on the GPU box the staged
copy ``oracle/_ref/`` (made by ``oracle/make_ref.py``, git-ignored) and the
committed fixtures under ``tests/golden/`` stand in for it.
"""
import importlib.util
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("SUBPIXAL_REFERENCE", "/root/reference")
# byte-for-byte staging copy made by oracle/make_ref.py (git-ignored; travels to the GPU box)
STAGED_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
_PKG = "subpixal"
_cache = {}


def root():
    """Directory holding ``subpixal/``: the mounted reference tree, else the staged copy."""
    for r in (REFERENCE_ROOT, STAGED_ROOT):
        if os.path.isfile(os.path.join(r, _PKG, "cc.py")):
            return r
    return None


def available():
    return root() is not None


def _bitfield_to_boolean_mask(bitfield, ignore_flags=0, flip_bits=None,
                              good_mask_value=True, dtype=np.bool_):
    """Minimal stand-in for stsci.tools.bitmask.bitfield_to_boolean_mask
    (un-vendored).  Semantics documented at cutout.py:77-109: ``ignore_flags``
    is the integer sum of "good" bits, ``None`` means ignore the DQ array."""
    bitfield = np.asarray(bitfield)
    if ignore_flags is None:
        bad = np.zeros(bitfield.shape, dtype=bool)
    else:
        ignore = int(ignore_flags)
        bad = (bitfield & ~np.array(ignore, dtype=bitfield.dtype)) != 0
    return (~bad if good_mask_value else bad).astype(dtype)


def _proj_plane_pixel_area(wcs):
    """Stand-in for astropy.wcs.utils.proj_plane_pixel_area (un-vendored; used by
    Cutout.pscale, cutout.py:1153-1157): |det| of the WCS's pixel scale matrix."""
    return abs(float(np.linalg.det(np.asarray(wcs.pixel_scale_matrix, dtype=np.float64))))


def _install_stubs():
    def mod(name, **attrs):
        m = sys.modules.get(name)
        if m is None:
            m = types.ModuleType(name)
            sys.modules[name] = m
        for k, v in attrs.items():
            setattr(m, k, v)
        return m

    if "astropy" not in sys.modules:
        a = mod("astropy")
        a.io = mod("astropy.io")
        a.io.fits = mod("astropy.io.fits")
        a.wcs = mod("astropy.wcs")
        a.wcs.utils = mod("astropy.wcs.utils", proj_plane_pixel_area=_proj_plane_pixel_area)
    if "stwcs" not in sys.modules:
        s = mod("stwcs")
        s.wcsutil = mod("stwcs.wcsutil", HSTWCS=object)
    if "stsci" not in sys.modules:
        s = mod("stsci")
        s.tools = mod("stsci.tools")
        s.tools.bitmask = mod(
            "stsci.tools.bitmask",
            bitfield_to_boolean_mask=_bitfield_to_boolean_mask)
    pkg = mod(_PKG, __version__="reference", __version_date__="n/a")
    pkg.__path__ = [os.path.join(root(), _PKG)]
    mod(_PKG + ".catalogs", ImageCatalog=type("ImageCatalog", (), {}))
    if not hasattr(np, "bool"):
        np.bool = bool  # cc.py:150


def load(name):
    """Return reference module ``subpixal.<name>`` (utils, centroid, cc,
    cutout), executed from its unmodified source file."""
    if name in _cache:
        return _cache[name]
    if not available():
        raise RuntimeError("reference tree not present at %s nor staged in %s"
                           % (REFERENCE_ROOT, STAGED_ROOT))
    _install_stubs()
    for dep in {"centroid": ["utils"], "cc": ["utils", "centroid"],
                "cutout": []}.get(name, []):
        load(dep)
    full = _PKG + "." + name
    spec = importlib.util.spec_from_file_location(
        full, os.path.join(root(), _PKG, name + ".py"))
    m = importlib.util.module_from_spec(spec)
    sys.modules[full] = m
    spec.loader.exec_module(m)
    _cache[name] = m
    return m
