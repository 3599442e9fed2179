"""CPU restatement of the FITS image decoding the reference gets from astropy.io.fits
(`h[ext].data`, align.py:311-320, 466-486; `fits.getdata`, catalogs.py:757).

TEST INFRASTRUCTURE ONLY: imported by tests/ - never by the product package.

astropy is an un-vendored, un-pinned dependency of the reference (setup.py:104-113) and is
absent from this image, so this follows its published behaviour (ImageHDU.data with the
defaults ``do_not_scale_image_data=False``, ``uint=True``: `_get_scaled_image_data`) and the
FITS standard 4.0 (big-endian two's-complement integers and IEEE-754 floats, sections
5.2 / 5.3; BSCALE / BZERO, section 4.4.2.5): **parity with astropy itself is unpinned**; the
byte order and integer / float conversions are numpy's own big-endian dtypes.
"""
import numpy as np

_RAW = {8: ">u1", 16: ">i2", 32: ">i4", 64: ">i8", -32: ">f4", -64: ">f8"}


def decode_payload(raw_bytes, bitpix, shape, bscale=1.0, bzero=0.0, uint=True, dtype=None):
    """What ``hdu.data`` holds (native byte order).  Unscaled payloads keep their type;
    BSCALE = 1 with BZERO = 2**(BITPIX-1) is the unsigned-integer convention (BITPIX 8 with
    BZERO -128: int8); otherwise the payload is converted to float32 (BITPIX 8 / 16 / -32)
    or float64 (32 / 64 / -64) and scaled IN PLACE: ``data *= bscale`` (if != 1), then
    ``data += bzero`` (if != 0).  ``dtype`` forces the floating type of the result."""
    n = int(np.prod(shape, dtype=np.int64))
    raw = np.frombuffer(raw_bytes, dtype=_RAW[bitpix], count=n).reshape(shape)
    if dtype is None and bscale == 1.0 and bzero == 0.0:
        return raw.astype(raw.dtype.newbyteorder("="))
    if dtype is None and uint and bitpix > 0 and bscale == 1.0:
        if bitpix == 8 and bzero == -128.0:
            return (raw.astype(np.int16) - 128).astype(np.int8)
        if bitpix > 8 and bzero == float(1 << (bitpix - 1)):
            unsigned = np.dtype("u%d" % (bitpix // 8))
            return (raw.astype(raw.dtype.newbyteorder("=")).view(unsigned)
                    ^ unsigned.type(1 << (bitpix - 1)))
    if dtype is None:
        dtype = np.float64 if abs(bitpix) > 16 and bitpix != -32 else np.float32
    data = np.array(raw, dtype=dtype)
    if bscale != 1.0:
        np.multiply(data, data.dtype.type(bscale), out=data)
    if bzero != 0.0:
        data += data.dtype.type(bzero)
    return data


def subtract_sky(data, sky):
    """``img_sci -= image_sky[ext]`` (align.py:480-481) in the array's own precision."""
    out = np.array(data)
    out -= out.dtype.type(sky)
    return out


def zero_plane(data, dtype=np.int64):
    """``np.equal(h[ext].data, 0.0).astype(np.int)`` (align.py:312)."""
    return np.equal(data, 0.0).astype(dtype)
