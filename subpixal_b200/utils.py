"""Scalar helpers kept on the host (reference utils.py).  Only ``py2round``
(reference utils.py:144-161) is on the hot path's public surface; file-name
parsing and temporary FITS helpers are out of scope."""
import numpy as np

__all__ = ["py2round", "native_array"]


def py2round(x):
    """Round half away from zero, like Python 2's ``round`` (utils.py:144-161).
    Scalars give a numpy float, iterables an array of the input dtype."""
    if hasattr(x, "__iter__"):
        arr = np.asarray(x)
        out = np.empty_like(arr)
        pos = arr >= 0.0
        out[pos] = np.floor(arr[pos] + 0.5)
        out[~pos] = np.ceil(arr[~pos] - 0.5)
        return out
    return np.floor(x + 0.5) if x >= 0.0 else np.ceil(x - 0.5)


def native_array(a, dtype=None):
    """C-contiguous array in NATIVE byte order.  Arrays read with astropy.io.fits are
    big-endian ('>f4', '>i2'), which ``torch.from_numpy`` rejects; the reference's
    numpy code takes them as they are (align.py:315-320, 476-483)."""
    a = np.asarray(a) if dtype is None else np.asarray(a, dtype=dtype)
    if not a.dtype.isnative:
        a = a.astype(a.dtype.newbyteorder('='))
    return np.ascontiguousarray(a)
