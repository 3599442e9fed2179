"""Image cutouts and the builders that map them between images.

Public surface of the reference ``subpixal.cutout`` (cutout.py:19-20): the
``Cutout`` class, ``create_primary_cutouts``, ``create_cutouts`` (+ the two
helpers it is made of) and the two overlap exceptions.  Box arithmetic and the
WCS calls (<= 5 points per cutout, astropy/wcslib) stay on the host; every
pixel operation - crop, fill, non-finite / DQ / segmentation masking - runs in
the ``spx_gather_cutouts`` CUDA kernel.  The builders issue ONE gather launch
per image for all cutouts instead of one numpy crop per source.
"""
import math

import numpy as np
import torch

from . import batch
from .utils import native_array

__all__ = ['Cutout', 'create_primary_cutouts', 'create_cutouts',
           'create_input_image_cutouts', 'drz_from_input_cutouts',
           'NoOverlapError', 'PartialOverlapError']

_FLOATS = (np.float32, np.float64)


class NoOverlapError(ValueError):
    """The extraction box lies entirely outside the image (cutout.py:37-39)."""


class PartialOverlapError(ValueError):
    """mode='strict' and the box sticks out of the image (cutout.py:42-44)."""


# --------------------------------------------------------------------------
# small host helpers
# --------------------------------------------------------------------------
def _ceil(v):
    return int(math.ceil(v))


def _floor(v):
    return int(math.floor(v))


def _round_pad(pad):
    return _ceil(pad) if pad >= 0 else _floor(pad)      # cutout.py:136, 391


def interpret_bit_flags(flags):
    """DQ "good bits" specification -> integer or None (documented at
    cutout.py:77-109; the reference delegates to stsci.tools.bitmask).
    Accepts None, an int (negative = two's-complement inverse), or a string
    such as ``'4,8'``, ``'4+8'``, ``'~4+8'``, ``'~0'``."""
    if flags is None:
        return None
    if isinstance(flags, str):
        txt = flags.strip()
        if txt.upper() in ('', 'NONE', 'INDEF'):
            return None
        flip = txt.startswith('~')
        if flip:
            txt = txt[1:].strip()
        if txt.startswith('(') and txt.endswith(')'):
            txt = txt[1:-1]
        if ',' in txt and '+' in txt:
            raise ValueError("Only one type of bit flag separator may be used.")
        parts = [p for p in txt.replace('+', ',').split(',') if p.strip() != '']
        total = 0
        for p in parts:
            total |= int(p)
        return ~total if flip else total
    return int(flags)


def _dq_device_view(arr):
    """numpy DQ array -> (same-width signed view for torch, original dtype).
    Big-endian arrays (FITS) are converted to native byte order first."""
    a = native_array(arr)
    if a.dtype.kind not in 'iub':
        raise TypeError("DQ array must have an integer dtype")
    signed = {1: np.int8, 2: np.int16, 4: np.int32, 8: np.int64}[a.dtype.itemsize]
    return a.view(signed), a.dtype


def _device_frame(data):
    """numpy image -> (native float32/float64 array to upload, dtype to restore or None).
    The reference crops any dtype (cutout.py:737-752: ``np.full(..., dtype=data.dtype)``);
    the gather kernel reads float32 / float64, so integer, bool and half-precision
    frames travel as float64 / float32 (exact) and the crops are cast back."""
    orig = np.asarray(data).dtype
    a = native_array(data)
    if a.dtype.type in _FLOATS:
        return a, (None if orig.isnative else orig)     # crops keep the input dtype
    if a.dtype.kind in 'iub':
        if a.dtype.itemsize >= 8 and a.size and \
                max(abs(int(a.max())), abs(int(a.min()))) > (1 << 53):
            raise TypeError("subpixal_b200: 64-bit integer image data beyond 2**53 cannot "
                            "be cropped exactly")
        return a.astype(np.float64), orig
    if a.dtype == np.float16:
        return a.astype(np.float32), orig
    raise TypeError("subpixal_b200 cutouts support real floating-point and integer image "
                    "data, got %s" % a.dtype)


def _box_checks(nx, ny, blc, trc, mode):
    """Argument checks of Cutout.__init__ (cutout.py:713-729)."""
    if blc[0] >= nx or blc[1] >= ny or trc[0] < 0 or trc[1] < 0:
        raise NoOverlapError("Cutout's extraction box does not overlap image "
                             "data array.")
    if trc[0] < blc[0] or trc[1] < blc[1]:
        raise ValueError("Ill-formed extraction box: coordinates of the "
                         "top-right corner cannot be smaller than the "
                         "coordinates of the bottom-left corner.")
    if mode == 'strict' and (blc[0] < 0 or blc[1] < 0 or
                             trc[0] >= nx or trc[1] >= ny):
        raise PartialOverlapError("Cutout's extraction box only partially "
                                  "overlaps image data array.")


def _overlap_slices(nx, ny, blc, trc):
    """(extraction slice in the image, insertion slice in the cutout),
    cutout.py:742-747."""
    bx1, by1 = max(0, blc[0]), max(0, blc[1])
    bx2, by2 = min(nx - 1, trc[0]) + 1, min(ny - 1, trc[1]) + 1
    return (np.s_[by1:by2, bx1:bx2],
            np.s_[by1 - blc[1]:by2 - blc[1], bx1 - blc[0]:bx2 - blc[0]])


_astropy_pixel_area = False          # False: not looked for yet; None: astropy absent


def _proj_plane_pixel_area():
    """``astropy.wcs.utils.proj_plane_pixel_area`` or None, looked up once (a failing import
    is retried - and paid for - on every call otherwise)."""
    global _astropy_pixel_area
    if _astropy_pixel_area is False:
        try:
            from astropy.wcs.utils import proj_plane_pixel_area
            _astropy_pixel_area = proj_plane_pixel_area
        except ImportError:
            _astropy_pixel_area = None
    return _astropy_pixel_area


class _DeviceImage(object):
    """An image (+ optional dq / weight / segmentation) resident on the GPU."""

    def __init__(self, data, dq=None, weight=None, seg=None):
        batch.require_cuda()
        host, self.restore = _device_frame(data)
        self.shape = data.shape
        self.data = torch.from_numpy(host).cuda()
        self.dq = self.dq_dtype = self.weight = self.seg = None
        if dq is not None:
            if dq.shape != data.shape:
                raise ValueError("Image's DQ array shape must match the shape "
                                 "of image 'data'.")
            view, self.dq_dtype = _dq_device_view(dq)
            self.dq = torch.from_numpy(view).cuda()
        if weight is not None:
            if weight.shape != data.shape:
                raise ValueError("Image's weight array shape must match the "
                                 "shape of image 'data'.")
            w = native_array(weight)
            if w.dtype.type not in _FLOATS:
                w = w.astype(np.float64)
            self.weight = torch.from_numpy(w).cuda()
        if seg is not None:
            s = native_array(seg)
            if s.dtype not in (np.int32, np.int64):
                s = s.astype(np.int64)
            self.seg = torch.from_numpy(s).cuda()

    def gather(self, boxes, fillval, src_ids=None, dq_good_bits=None):
        """One launch for all boxes; returns per-cutout numpy arrays."""
        if self.restore is not None:
            # what numpy stores for ``fillval`` in this dtype (NaN in an integer array
            # raises here exactly as np.full does in the reference, cutout.py:737)
            fillval = np.full((), fillval, dtype=self.restore).astype(np.float64)
        res = batch.gather_cutouts(
            self.data, boxes, dq=self.dq, weight=self.weight,
            seg=self.seg if src_ids is not None else None, src_ids=src_ids,
            dq_good_bits=dq_good_bits, fillval=fillval)
        off = res["offsets"].numpy()
        shp = res["shapes"].numpy()
        flat = {k: res[k].cpu().numpy() for k in ("data", "mask", "dq", "weight")
                if res[k] is not None}
        out = []
        for k in range(len(shp)):
            h, w = int(shp[k, 0]), int(shp[k, 1])
            sl = slice(int(off[k]), int(off[k + 1]))
            item = {"data": (flat["data"][sl].reshape(h, w).copy() if self.restore is None
                             else flat["data"][sl].reshape(h, w).astype(self.restore)),
                    "mask": flat["mask"][sl].reshape(h, w).astype(np.bool_),
                    "dq": None, "weight": None}
            if "dq" in flat:
                item["dq"] = flat["dq"][sl].reshape(h, w).view(self.dq_dtype).copy()
            if "weight" in flat:
                item["weight"] = flat["weight"][sl].reshape(h, w).copy()
            out.append(item)
        return out


# --------------------------------------------------------------------------
# Cutout
# --------------------------------------------------------------------------
class Cutout(object):
    """A rectangular crop of an image together with its bad-pixel mask, its
    DQ / weight crops and what is needed to map cutout pixels to the sky.

    Parameters (reference cutout.py:689-691)
    ----------
    data : numpy.ndarray or None
        Image to crop.  With ``None`` an empty cutout of the requested size is
        made (``trc`` is then mandatory).
    wcs : astropy.wcs.WCS-like or None
        Kept by reference; used only by the WCS methods.
    blc, trc : (x, y)
        Bottom-left / top-right pixel of the box; ``trc`` is inclusive and
        defaults to the last pixel of ``data``.
    src_pos : (x, y), optional
        Source position in IMAGE coordinates (default: box centre).
    src_weight, src_id, data_units ('counts' | 'rate'), exptime : bookkeeping.
    dq, weight : numpy.ndarray, optional
        Cropped the same way, zero-filled outside the image.
    mode : 'strict' | 'fill'
        'strict' raises `PartialOverlapError` when the box leaves the image;
        'fill' pads the outside with ``fillval``.

    Attributes: ``data mask dq weight blc trc dx dy width height naxis src_pos
    cutout_src_pos src_id src_weight exptime data_units extraction_slice
    insertion_slice wcs pscale``; ``mask`` is True for BAD pixels (outside the
    image or non-finite).
    """
    DEFAULT_ACCURACY = 1.0e-5
    DEFAULT_MAXITER = 50
    DEFAULT_QUIET = True

    def __init__(self, data, wcs, blc=(0, 0), trc=None,
                 src_pos=None, src_weight=None, dq=None, weight=None, src_id=0,
                 data_units='rate', exptime=1, mode='strict', fillval=np.nan):
        if data is None and trc is None:
            raise ValueError("'trc' cannot be None when 'data' is None.")
        if mode not in ('strict', 'fill'):
            raise ValueError("Argument 'mode' must be either 'strict' or "
                             "'fill'.")
        if data is None:
            nx, ny = trc[0] + 1, trc[1] + 1
        else:
            data = np.asarray(data)
            ny, nx = data.shape
        if trc is None:
            trc = (nx - 1, ny - 1)
        _box_checks(nx, ny, blc, trc, mode)
        blc, trc = (int(blc[0]), int(blc[1])), (int(trc[0]), int(trc[1]))
        ex, ins = _overlap_slices(nx, ny, blc, trc)

        if data is None:
            # nothing to read: an empty float32 canvas (cutout.py:697-700)
            h, w = trc[1] - blc[1] + 1, trc[0] - blc[0] + 1
            arrays = {"data": np.full((h, w), fillval, dtype=np.float32),
                      "dq": None, "weight": None}
            bad = np.ones((h, w), dtype=np.bool_)
            bad[ins] = False
            arrays["mask"] = bad | ~np.isfinite(arrays["data"])
            for name, arr in (("dq", dq), ("weight", weight)):
                if arr is not None:
                    raise ValueError("'%s' requires image 'data'." % name)
        else:
            for name, arr in (("DQ", dq), ("weight", weight)):
                if arr is not None and arr.shape != (ny, nx):
                    raise ValueError("Image's %s array shape must match the "
                                     "shape of image 'data'." % name)
            # upload only the overlap; the kernel crops/fills/masks it
            sub = _DeviceImage(data[ex],
                               None if dq is None else np.asarray(dq)[ex],
                               None if weight is None else np.asarray(weight)[ex])
            x0, y0 = ex[1].start, ex[0].start
            arrays = sub.gather([[blc[0] - x0, blc[1] - y0, trc[0] - x0, trc[1] - y0]],
                                fillval)[0]
            if weight is not None and arrays["weight"].dtype != weight.dtype:
                arrays["weight"] = arrays["weight"].astype(weight.dtype)
        self._assemble(arrays, wcs, blc, trc, (nx, ny), ex, ins, src_pos,
                       src_weight, src_id, data_units, exptime)

    # shared tail of __init__ and of the batched builders
    def _assemble(self, arrays, wcs, blc, trc, onxy, ex, ins, src_pos,
                  src_weight, src_id, data_units, exptime):
        self._onx, self._ony = onxy
        self._blc, self._trc = tuple(blc), tuple(trc)
        self.src_pos = src_pos
        self.src_weight = src_weight
        self._data = arrays["data"]
        self._mask = arrays["mask"]
        self._dq = arrays["dq"]
        self._weight = arrays["weight"]
        self._src_id = src_id
        self._dx = 0
        self._dy = 0
        self._wcs = wcs
        self._naxis = [self.width, self.height]
        self.naxis1, self.naxis2 = self._naxis
        self._eslice, self._islice = ex, ins
        self.exptime = exptime
        self.data_units = data_units

    def __deepcopy__(self, memo):
        """``copy.deepcopy`` of a cutout (blot.py:152, align.py's histories): arrays are copied,
        immutable attributes (numbers, strings, tuples of numbers, slices) are shared, anything
        else - the WCS object, unless the caller put it into ``memo`` - goes through
        ``copy.deepcopy``.  Same result as the generic protocol at a tenth of its cost (it walks
        about 70 objects per cutout)."""
        import copy
        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        flat = (int, float, complex, str, bytes, bool, type(None), slice, np.generic)
        for key, val in self.__dict__.items():
            if isinstance(val, np.ndarray):
                new.__dict__[key] = val.copy()
            elif isinstance(val, flat) or (isinstance(val, tuple) and
                                           all(isinstance(v, flat) for v in val)):
                new.__dict__[key] = val
            else:
                new.__dict__[key] = copy.deepcopy(val, memo)
        return new

    @classmethod
    def _from_arrays(cls, arrays, wcs, blc, trc, onxy, **meta):
        self = cls.__new__(cls)
        ex, ins = _overlap_slices(onxy[0], onxy[1], blc, trc)
        self._assemble(arrays, wcs, blc, trc, onxy, ex, ins,
                       meta.get("src_pos"), meta.get("src_weight"),
                       meta.get("src_id", 0), meta.get("data_units", 'rate'),
                       meta.get("exptime", 1))
        return self

    # ---- geometry ---------------------------------------------------------
    @property
    def blc(self):
        """(x, y) of the bottom-left pixel in the parent image."""
        return self._blc

    @blc.setter
    def blc(self, value):
        self._blc = (value[0], value[1])

    @property
    def trc(self):
        """(x, y) of the top-right pixel (inclusive) in the parent image."""
        return self._trc

    @trc.setter
    def trc(self, value):
        self._trc = (value[0], value[1])

    @property
    def width(self):
        return self._trc[0] - self._blc[0] + 1

    @property
    def height(self):
        return self._trc[1] - self._blc[1] + 1

    @property
    def naxis(self):
        """[width, height] - FITS NAXIS order."""
        return self._naxis

    @property
    def dx(self):
        """Displacement of the cutout grid along x (pixels); `blot` shifts it
        by half a pixel to make the four phases (align.py:667-679)."""
        return self._dx

    @dx.setter
    def dx(self, value):
        self._dx = value

    @property
    def dy(self):
        return self._dy

    @dy.setter
    def dy(self, value):
        self._dy = value

    @property
    def extraction_slice(self):
        """Slice of the parent image that was copied."""
        return self._eslice

    @property
    def insertion_slice(self):
        """Slice of the cutout array that received image pixels."""
        return self._islice

    # ---- source bookkeeping ----------------------------------------------
    def _centre(self):
        return (0.5 * (self._blc[0] + self._trc[0]),
                0.5 * (self._blc[1] + self._trc[1]))

    @property
    def src_pos(self):
        """Source position in parent-image pixels."""
        return self._src_pos

    @src_pos.setter
    def src_pos(self, pos):
        self._src_pos = self._centre() if pos is None else tuple(pos)[:2]

    @property
    def cutout_src_pos(self):
        """Source position in cutout pixels."""
        return (self._src_pos[0] - self._blc[0], self._src_pos[1] - self._blc[1])

    @cutout_src_pos.setter
    def cutout_src_pos(self, pos):
        if pos is None:
            self._src_pos = self._centre()
        else:
            self._src_pos = (pos[0] + self._blc[0], pos[1] + self._blc[1])

    @property
    def src_weight(self):
        return self._src_weight

    @src_weight.setter
    def src_weight(self, value):
        if value is not None and np.any(value < 0.0):
            raise ValueError("Source weight must be a non-negative number "
                             "or None.")
        self._src_weight = value

    @property
    def src_id(self):
        return self._src_id

    @src_id.setter
    def src_id(self, value):
        self._src_id = value

    @property
    def exptime(self):
        return self._exptime

    @exptime.setter
    def exptime(self, value):
        if value <= 0:
            raise ValueError("'exptime' must be positive.")
        self._exptime = value

    @property
    def data_units(self):
        """'rate' or 'counts'."""
        return self._data_units

    @data_units.setter
    def data_units(self, units):
        units = units.lower()
        if units not in ('rate', 'counts'):
            raise ValueError("Allowed image data units are: 'rate' or "
                             "'counts'.")
        self._data_units = units

    # ---- pixel arrays -----------------------------------------------------
    def _same_shape(self, arr, what):
        if arr.shape != self._data.shape:
            raise ValueError(
                "ValueError: could not broadcast input array from shape "
                "({:s}) into shape ({:s})".format(
                    ','.join(map(str, arr.shape)),
                    ','.join(map(str, self._data.shape))) if what in ('data', 'mask')
                else "Cutout's {} array shape must match the shape of "
                     "cutout's image_data.".format(what))
        return arr

    @property
    def data(self):
        """Cutout pixels (writable numpy array)."""
        return self._data

    @data.setter
    def data(self, value):
        if value is None:
            raise ValueError("'data' cannot be None.")
        if value is not self._data:
            self._data = self._same_shape(np.array(value), 'data')

    @property
    def mask(self):
        """Boolean array, True where the pixel must not be used."""
        return self._mask

    @mask.setter
    def mask(self, value):
        if value is None:
            raise ValueError("'mask' cannot be None.")
        if value is not self._mask:
            self._mask = self._same_shape(np.array(value, dtype=np.bool_), 'mask')

    @property
    def dq(self):
        return self._dq

    @dq.setter
    def dq(self, value):
        if value is None:
            self._dq = None
        elif value is not self._dq:
            self._dq = self._same_shape(np.array(value), 'DQ')

    @property
    def weight(self):
        return self._weight

    @weight.setter
    def weight(self, value):
        if value is None:
            self._weight = None
        elif value is not self._weight:
            self._weight = self._same_shape(np.array(value), 'weight')

    # ---- coordinates (WCS work stays in astropy / wcslib) -----------------
    def get_bbox(self, wrt='orig'):
        """Outer corners of the box as a (4, 2) array, pixel centres being
        integers.  ``wrt``: 'orig' (parent image), 'blc' (cutout) or 'world'."""
        if wrt not in ('orig', 'blc', 'world'):
            raise ValueError("'wrt' must be one of the following: "
                             "'orig', 'blc', 'world'.")
        if wrt == 'orig':
            (x1, y1), (x2, y2) = self._blc, self._trc
        else:
            x1 = y1 = 0
            x2, y2 = self._trc[0] - self._blc[0], self._trc[1] - self._blc[1]
        box = np.asarray([[x1 - 0.5, y1 - 0.5], [x1 - 0.5, y2 + 0.5],
                          [x2 + 0.5, y2 + 0.5], [x2 + 0.5, y1 - 0.5]])
        return self.pix2world(box) if wrt == 'world' else box

    @staticmethod
    def _split_args(args, kind):
        if len(args) == 2:
            try:
                return (np.asarray(args[0], dtype=np.float64),
                        np.asarray(args[1], dtype=np.float64), True)
            except Exception:
                raise TypeError("When providing two arguments, they must be "
                                "(x, y) where x and y are Nx1 vectors.")
        if len(args) == 1:
            try:
                pair = np.asarray(args[0], dtype=np.float64)
                return pair[:, 0], pair[:, 1], False
            except Exception:
                raise TypeError("When providing one argument, it must be an "
                                "array of shape Nx2 containing %s." % kind)
        raise TypeError("Expected 2 or 3 arguments, {:d} given.".format(len(args)))

    def world2pix(self, *args, origin=0):
        """World -> CUTOUT pixel coordinates (cutout.py:1070-1110)."""
        ra, dec, vectors = self._split_args(args, "Ra & Dec")
        x, y = self._wcs.all_world2pix(ra, dec, origin,
                                       accuracy=self.DEFAULT_ACCURACY,
                                       maxiter=self.DEFAULT_MAXITER,
                                       quiet=self.DEFAULT_QUIET)
        x = x - (self._blc[0] - self._dx)
        y = y - (self._blc[1] - self._dy)
        return [x, y] if vectors else np.dstack([x, y])[0]

    def pix2world(self, *args, origin=0):
        """CUTOUT pixel -> world coordinates (cutout.py:1112-1145)."""
        x, y, vectors = self._split_args(args, "x & y")
        x = x + (self._blc[0] - self._dx)
        y = y + (self._blc[1] - self._dy)
        ra, dec = self._wcs.all_pix2world(x, y, origin)
        return [ra, dec] if vectors else np.dstack([ra, dec])[0]

    @property
    def wcs(self):
        return self._wcs

    @property
    def pscale(self):
        """Pixel scale in the tangent plane at the reference point."""
        if self._wcs is None:
            raise ValueError("WCS was not set. Unable to compute pixel scale.")
        area = _proj_plane_pixel_area()
        if area is not None:
            return np.sqrt(area(self._wcs))
        return math.sqrt(abs(np.linalg.det(np.asarray(self._wcs.pixel_scale_matrix))))


# --------------------------------------------------------------------------
# builders
# --------------------------------------------------------------------------
def _table_columns(catalog):
    """Accept an ImageCatalog, an astropy Table, a structured array or a dict
    of columns."""
    if hasattr(catalog, 'catalog') and not hasattr(catalog, 'colnames'):
        catalog = catalog.catalog                          # cutout.py:130-131
    if hasattr(catalog, 'colnames'):
        names = list(catalog.colnames)
    elif hasattr(catalog, 'dtype') and catalog.dtype.names:
        names = list(catalog.dtype.names)
    else:
        names = list(catalog.keys())
    return catalog, names


def _segment_boxes(dev_seg):
    """Bounding box of every label of a segmentation image in ONE GPU pass
    (``spx_segment_boxes``; the reference scans the whole frame once per source,
    cutout.py:151-162).  Returns labels and x1, x2, y1, y2 arrays (host)."""
    tab = batch.segment_boxes(dev_seg).cpu().numpy()
    labels = np.flatnonzero(tab[:, 1] >= 0)
    return labels, tab[labels, 0], tab[labels, 1], tab[labels, 2], tab[labels, 3]


def create_primary_cutouts(catalog, segmentation_image, imdata, imwcs,
                           imdq=None, dqbitmask=0, imweight=None,
                           data_units='counts', exptime=1, pad=1):
    """First-order cutouts from a (drizzle-)combined image: one per catalog
    source present in the segmentation image, boxed by the segment's bounding
    rectangle grown by ``pad``; pixels outside the segment or with bad DQ bits
    are masked; segments touching the frame edge and fully masked cutouts are
    dropped (reference cutout.py:47-198).  ``dqbitmask`` is the sum of the
    "good" DQ bits (None ignores DQ; see `interpret_bit_flags`)."""
    catalog, names = _table_columns(catalog)
    seg = np.asarray(segmentation_image)
    ny, nx = seg.shape
    pad = _round_pad(pad)
    if seg.shape != np.shape(imdata):
        raise ValueError("segmentation image and image data must have the same shape")
    dev = _DeviceImage(np.asarray(imdata), imdq, imweight, seg)
    labels, lx1, lx2, ly1, ly2 = _segment_boxes(dev.seg)
    ids, cat_idx, lab_idx = np.intersect1d(np.asarray(catalog['id']), labels,
                                           return_indices=True)   # cutout.py:139-143
    weights = catalog['weight'] if 'weight' in names else None
    good_bits = interpret_bit_flags(dqbitmask)

    jobs = []
    for sid, ci, li in zip(ids, cat_idx, lab_idx):
        x1, x2, y1, y2 = int(lx1[li]), int(lx2[li]), int(ly1[li]), int(ly2[li])
        if x1 <= 0 or y1 <= 0 or x2 >= nx - 1 or y2 >= ny - 1:   # cutout.py:164-169
            continue
        blc, trc = (x1 - pad, y1 - pad), (x2 + pad, y2 + pad)
        _box_checks(imdata.shape[1], imdata.shape[0], blc, trc, 'strict')
        jobs.append((sid, blc, trc, (catalog['x'][ci], catalog['y'][ci]),
                     None if weights is None else weights[ci]))
    if not jobs:
        return []
    crops = dev.gather([[b[0], b[1], t[0], t[1]] for _, b, t, _, _ in jobs], 0,
                       src_ids=[j[0] for j in jobs],
                       dq_good_bits=good_bits if imdq is not None else None)
    out = []
    for (sid, blc, trc, pos, wgt), arrays in zip(jobs, crops):
        if np.all(arrays["mask"]):                               # cutout.py:195
            continue
        out.append(Cutout._from_arrays(
            arrays, imwcs, blc, trc, (imdata.shape[1], imdata.shape[0]),
            src_pos=pos, src_weight=wgt, src_id=sid, data_units=data_units,
            exptime=exptime))
    return out


def _by_wcs(cutouts):
    """Indices of the cutouts grouped by their WCS OBJECT (the cutouts of one image share it):
    the builders make one WCS call per group instead of one or two per cutout."""
    groups = {}
    for k, ct in enumerate(cutouts):
        groups.setdefault(id(ct.wcs), []).append(k)
    return list(groups.values())


def _footprint_boxes(cutouts, imwcs, pad):
    """Conservative integer rectangle around every cutout's footprint in another image
    (cutout.py:265-272: ``imwcs.all_world2pix(ct.get_bbox('world'), 0, accuracy=1e-5,
    maxiter=50)``, floor / ceil of the extremes -/+ pad), with the four corners of all
    cutouts of a WCS group in ONE pix2world and ONE world2pix call - the same element-wise
    arithmetic as ``Cutout.get_bbox`` / ``Cutout.pix2world``."""
    boxes = [None] * len(cutouts)
    for idx in _by_wcs(cutouts):
        cts = [cutouts[k] for k in idx]
        m = len(cts)
        x2 = np.array([ct.trc[0] - ct.blc[0] for ct in cts], dtype=np.float64)
        y2 = np.array([ct.trc[1] - ct.blc[1] for ct in cts], dtype=np.float64)
        off = np.array([[ct.blc[0] - ct.dx, ct.blc[1] - ct.dy] for ct in cts], dtype=np.float64)
        zero = np.zeros(m)
        cx = np.stack([zero - 0.5, zero - 0.5, x2 + 0.5, x2 + 0.5], axis=1)      # get_bbox order
        cy = np.stack([zero - 0.5, y2 + 0.5, y2 + 0.5, zero - 0.5], axis=1)
        ra, dec = cts[0].wcs.all_pix2world((cx + off[:, 0:1]).ravel(), (cy + off[:, 1:2]).ravel(), 0)
        fp = np.asarray(imwcs.all_world2pix(np.dstack([ra, dec])[0], 0, accuracy=1e-5, maxiter=50))
        fx, fy = fp[:, 0].reshape(m, 4), fp[:, 1].reshape(m, 4)
        for j, k in enumerate(idx):
            boxes[k] = (_floor(fx[j].min() - pad), _floor(fy[j].min() - pad),
                        _ceil(fx[j].max() + pad), _ceil(fy[j].max() + pad))
    return boxes


def _finish_mapped(jobs, crops, imwcs, onxy, data_units, exptime):
    cutouts, sources = [], []
    for (ct, blc, trc), arrays in zip(jobs, crops):
        if np.all(arrays["mask"]):
            continue
        new = Cutout._from_arrays(arrays, imwcs, blc, trc, onxy,
                                  src_weight=ct.src_weight, src_id=ct.src_id,
                                  data_units=data_units, exptime=exptime)
        cutouts.append(new)
        sources.append(ct)
    # full WCS round trip for the source positions (cutout.py:300-301:
    # ``new.world2pix(ct.pix2world([ct.cutout_src_pos]))``), one call pair per WCS group
    for idx in _by_wcs(sources):
        cts = [sources[k] for k in idx]
        pos = np.array([ct.cutout_src_pos for ct in cts], dtype=np.float64).reshape(-1, 2)
        off = np.array([[ct.blc[0] - ct.dx, ct.blc[1] - ct.dy] for ct in cts], dtype=np.float64)
        ra, dec = cts[0].wcs.all_pix2world(pos[:, 0] + off[:, 0], pos[:, 1] + off[:, 1], 0)
        first = cutouts[idx[0]]
        x, y = imwcs.all_world2pix(np.asarray(ra, np.float64), np.asarray(dec, np.float64), 0,
                                   accuracy=first.DEFAULT_ACCURACY, maxiter=first.DEFAULT_MAXITER,
                                   quiet=first.DEFAULT_QUIET)
        for j, k in enumerate(idx):
            new = cutouts[k]
            new.cutout_src_pos = [float(x[j] - (new.blc[0] - new.dx)),
                                  float(y[j] - (new.blc[1] - new.dy))]
    return cutouts, sources


def create_input_image_cutouts(primary_cutouts, imdata, imwcs, imdq=None,
                               dqbitmask=0, imweight=None, data_units='counts',
                               exptime=1, pad=1):
    """Map cutouts of one image onto another image (typically drizzled ->
    distorted input).  A cutout whose padded footprint does not fit inside
    ``imdata`` is dropped.  Returns ``(new_cutouts, matching_input_cutouts)``
    (reference cutout.py:201-306)."""
    imdata = np.asarray(imdata)
    ny, nx = imdata.shape
    jobs = []
    primary_cutouts = list(primary_cutouts)
    for ct, (x1, y1, x2, y2) in zip(primary_cutouts, _footprint_boxes(primary_cutouts, imwcs, pad)):
        if (x1 < 0 or x1 >= nx or y1 < 0 or y1 > ny or
                x2 < 0 or x2 >= nx or y2 < 0 or y2 > ny):        # cutout.py:276-278
            continue
        try:
            _box_checks(nx, ny, (x1, y1), (x2, y2), 'strict')
        except (NoOverlapError, PartialOverlapError):            # cutout.py:294-295
            continue
        jobs.append((ct, (x1, y1), (x2, y2)))
    if not jobs:
        return [], []
    dev = _DeviceImage(imdata, imdq, imweight)
    crops = dev.gather([[b[0], b[1], t[0], t[1]] for _, b, t in jobs], 0,
                       dq_good_bits=(interpret_bit_flags(dqbitmask)
                                     if imdq is not None else None))
    return _finish_mapped(jobs, crops, imwcs, (nx, ny), data_units, exptime)


def drz_from_input_cutouts(input_cutouts, segmentation_image, imdata, imwcs,
                           imdq=None, dqbitmask=0, imweight=None,
                           data_units='counts', exptime=1, pad=1,
                           combine_seg_mask=True):
    """Map cutouts back onto a (drizzled) image in 'fill' mode: partial
    overlaps are zero-filled, pixels outside the source's segment are masked
    when ``combine_seg_mask`` is set (reference cutout.py:309-438)."""
    imdata = np.asarray(imdata)
    ny, nx = imdata.shape
    pad = _round_pad(pad)
    jobs = []
    input_cutouts = list(input_cutouts)
    for ct, (x1, y1, x2, y2) in zip(input_cutouts, _footprint_boxes(input_cutouts, imwcs, pad)):
        try:
            _box_checks(nx, ny, (x1, y1), (x2, y2), 'fill')
        except NoOverlapError:                                   # cutout.py:412-413
            continue
        jobs.append((ct, (x1, y1), (x2, y2)))
    if not jobs:
        return [], []
    dev = _DeviceImage(imdata, imdq, imweight,
                       np.asarray(segmentation_image) if combine_seg_mask else None)
    crops = dev.gather([[b[0], b[1], t[0], t[1]] for _, b, t in jobs], 0,
                       src_ids=[j[0].src_id for j in jobs] if combine_seg_mask else None,
                       dq_good_bits=(interpret_bit_flags(dqbitmask)
                                     if imdq is not None else None))
    return _finish_mapped(jobs, crops, imwcs, (nx, ny), data_units, exptime)


def create_cutouts(primary_cutouts, segmentation_image,
                   drz_data, drz_wcs, flt_data, flt_wcs,
                   drz_dq=None, drz_dqbitmask=0, drz_weight=None,
                   drz_data_units='rate', drz_exptime=1,
                   flt_dq=None, flt_dqbitmask=0, flt_weight=None,
                   flt_data_units='counts', flt_exptime=1,
                   pad=2, combine_seg_mask=True):
    """primary cutouts -> cutouts in the distorted ("flt") image -> cutouts in
    a drizzled image.  Returns the aligned lists ``(flt_cutouts, drz_cutouts)``,
    or ``([], [])`` (reference cutout.py:441-587)."""
    flt_first, primaries = create_input_image_cutouts(
        primary_cutouts, flt_data, flt_wcs, imdq=flt_dq, dqbitmask=flt_dqbitmask,
        imweight=flt_weight, data_units=flt_data_units, exptime=flt_exptime,
        pad=pad)
    if not flt_first:
        return [], []
    ratio = flt_first[0].pscale / primaries[0].pscale                    # cutout.py:570
    drz_cutouts, flt_cutouts = drz_from_input_cutouts(
        flt_first, segmentation_image, drz_data, drz_wcs, imdq=drz_dq,
        dqbitmask=drz_dqbitmask, imweight=drz_weight,
        data_units=drz_data_units, exptime=drz_exptime,
        pad=pad * ratio, combine_seg_mask=combine_seg_mask)
    return flt_cutouts, drz_cutouts
