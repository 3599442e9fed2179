"""subpixal_b200 - B200-native (sm_100a) implementation of subpixal's hot path.

Drop-in entry points with the reference's signatures:

* ``subpixal_b200.cc.find_displacement``      (reference cc.py:21)
* ``subpixal_b200.centroid.find_peak``        (reference centroid.py:18)
* ``subpixal_b200.cutout.Cutout`` / ``create_primary_cutouts`` /
  ``create_cutouts`` / ``NoOverlapError`` / ``PartialOverlapError``
                                              (reference cutout.py:19-20)
* ``subpixal_b200.utils.py2round``            (reference utils.py:144)

plus the batched tensor API in ``subpixal_b200.batch`` (what a batched
``find_linear_fit`` calls) and the multi-GPU sharding in ``subpixal_b200.dist``.
All arithmetic runs in hand-written CUDA kernels behind the C ABI declared in
``include/subpixal_b200.h``; there is no CPU fallback.
"""
from .utils import py2round                                   # noqa: F401
from .centroid import find_peak                               # noqa: F401
from .cc import find_displacement                             # noqa: F401
from .cutout import (Cutout, create_primary_cutouts, create_cutouts,  # noqa: F401
                     create_input_image_cutouts, drz_from_input_cutouts,
                     NoOverlapError, PartialOverlapError)

__version__ = "0.1.0"
__all__ = ["find_displacement", "find_peak", "Cutout", "create_primary_cutouts",
           "create_cutouts", "NoOverlapError", "PartialOverlapError", "py2round"]
