"""ctypes binding of the C-ABI CUDA library (include/subpixal_b200.h).

There is NO fallback: if the shared object is missing, or no CUDA device is
visible when a compute entry point is called, an exception is raised.
"""
import ctypes as C
import os

from .build import LIB_PATH

# ---- constants mirrored from include/subpixal_b200.h -------------------------
ABI_VERSION = 1
E_BADARG, E_UNSUPPORTED, E_WORKSPACE, E_DTYPE = -1, -2, -3, -4
MAX_CUTOUT, MIN_CUTOUT, MAX_FIT_BOX = 256, 3, 15
CC, NCC, ZNCC = 0, 1, 2
F_ZERO_BAD_PHASES, F_ZERO_BAD_REF, F_EXACT_NORM = 0x1, 0x2, 0x4
S_NONFINITE_REF = 0x001
S_NONFINITE_P00, S_NONFINITE_P10, S_NONFINITE_P01, S_NONFINITE_P11 = 0x2, 0x4, 0x8, 0x10
S_BAD_NORM, S_PEAK_EDGE, S_NO_MAXIMUM, S_FIT_OUTSIDE = 0x20, 0x40, 0x80, 0x100
S_FIT_NONFINITE, S_BOX_MISMATCH, S_FEW_POINTS, S_SINGULAR = 0x200, 0x400, 0x800, 0x1000
S_UNSUPPORTED_SIZE = 0x2000
T_F32, T_F64, T_I8, T_I16, T_I32, T_I64 = 0, 1, 2, 3, 4, 5

EXPORTS = ("spx_abi_version", "spx_error_string", "spx_gather_cutouts",
           "spx_cc_workspace_bytes", "spx_cc_fit_stack", "spx_cc_fit_ragged",
           "spx_cc_fit_frames",
           "spx_find_peak", "spx_segment_boxes", "spx_blot_phases",
           "spx_blot_workspace_bytes", "spx_fits_decode")
BLOT_NEAREST, BLOT_LINEAR, BLOT_POLY5, BLOT_POLY3 = 0, 1, 2, 3
BLOT_MAX_PHASES = 8


class GatherArgs(C.Structure):
    """struct spx_gather_args"""
    _fields_ = [
        ("frame", C.c_void_p), ("frame_kind", C.c_int),
        ("frame_h", C.c_int), ("frame_w", C.c_int), ("frame_pitch", C.c_int64),
        ("dq", C.c_void_p), ("dq_kind", C.c_int),
        ("weight", C.c_void_p), ("weight_kind", C.c_int),
        ("seg", C.c_void_p), ("seg_kind", C.c_int),
        ("src_ids", C.c_void_p),
        ("dq_use_bits", C.c_int), ("dq_good_bits", C.c_int64),
        ("boxes", C.c_void_p), ("out_offset", C.c_void_p), ("n", C.c_int),
        ("max_h", C.c_int), ("max_w", C.c_int),
        ("fillval", C.c_double), ("zero_masked", C.c_int),
        ("data_out", C.c_void_p), ("mask_out", C.c_void_p),
        ("dq_out", C.c_void_p), ("weight_out", C.c_void_p),
    ]


class BlotArgs(C.Structure):
    """struct spx_blot_args"""
    _fields_ = [
        ("src", C.c_void_p), ("src_offset", C.c_void_p), ("src_shape", C.c_void_p),
        ("map", C.c_void_p), ("scale", C.c_void_p),
        ("tgt_shape", C.c_void_p), ("out_offset", C.c_void_p),
        ("out", C.c_void_p),
        ("phases_host", C.POINTER(C.c_double)),
        ("n", C.c_int), ("nph", C.c_int), ("interp", C.c_int),
        ("max_ny", C.c_int), ("max_nx", C.c_int),
        ("misval", C.c_float),
        ("workspace", C.c_void_p), ("ws_offset", C.c_void_p),
        ("max_src_h", C.c_int), ("max_src_w", C.c_int),
    ]


class SubpixalError(RuntimeError):
    """A C-ABI call returned a non-zero code."""


_lib = None


def load():
    """Load the shared library (once) and declare the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("SUBPIXAL_B200_LIB", LIB_PATH)   # developer override (kernel variants)
    if not os.path.exists(path):
        raise RuntimeError(
            "subpixal_b200: CUDA library %s is missing and there is no CPU "
            "fallback. Build it with `python -m subpixal_b200.build` "
            "(needs nvcc, targets sm_100a)." % path)
    lib = C.CDLL(path)
    vp, i32, i64, u32, sz = C.c_void_p, C.c_int, C.c_int64, C.c_uint, C.c_size_t
    lib.spx_abi_version.restype = i32
    lib.spx_abi_version.argtypes = []
    lib.spx_error_string.restype = C.c_char_p
    lib.spx_error_string.argtypes = [i32]
    lib.spx_gather_cutouts.restype = i32
    lib.spx_gather_cutouts.argtypes = [C.POINTER(GatherArgs), vp]
    lib.spx_cc_workspace_bytes.restype = sz
    lib.spx_cc_workspace_bytes.argtypes = [i32, i32, i32, i32]
    lib.spx_cc_fit_stack.restype = i32
    lib.spx_cc_fit_stack.argtypes = [vp, vp, vp, vp, vp, C.POINTER(i64 * 5),
                                     i32, i32, i32, i32, vp, vp, vp, vp, vp, sz, vp]
    lib.spx_cc_fit_ragged.restype = i32
    lib.spx_cc_fit_ragged.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32,
                                      vp, vp, vp, vp, vp, vp, sz, vp]
    lib.spx_cc_fit_frames.restype = i32
    lib.spx_cc_fit_frames.argtypes = [C.POINTER(vp * 5), i32, i32, i64, vp,
                                      i32, i32, i32, i32, u32,
                                      vp, vp, vp, vp, vp, sz, vp]
    lib.spx_find_peak.restype = i32
    lib.spx_find_peak.argtypes = [vp, i32, i32, i32, i32, vp, vp,
                                  i32, i32, i32, i32, vp, vp, vp, vp]
    lib.spx_segment_boxes.restype = i32
    lib.spx_segment_boxes.argtypes = [vp, i32, i32, i32, i64, i32, vp, vp]
    lib.spx_blot_workspace_bytes.restype = sz
    lib.spx_blot_workspace_bytes.argtypes = [i32, i32]
    lib.spx_blot_phases.restype = i32
    lib.spx_blot_phases.argtypes = [C.POINTER(BlotArgs), vp]
    lib.spx_fits_decode.restype = i32
    lib.spx_fits_decode.argtypes = [vp, i64, i32, C.c_double, C.c_double, C.c_double,
                                    i32, vp, i32, vp, vp]
    if lib.spx_abi_version() != ABI_VERSION:
        raise RuntimeError("subpixal_b200: ABI version mismatch, rebuild the library")
    _lib = lib
    return lib


def check(code):
    if code != 0:
        msg = load().spx_error_string(code).decode()
        raise SubpixalError("subpixal_b200 C-ABI error %d: %s" % (code, msg))
