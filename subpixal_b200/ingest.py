"""FITS / catalogue ingest straight into device memory.

What the reference does around the hot path with astropy (absent from this image):

* ``utils.parse_file_name`` (utils.py:21-93): ``'x_flt.fits[sci,2]'`` -> file + extension;
* ``fits.open(fn)[ext].data`` / header keywords ``EXPTIME``, ``BUNIT`` for the drizzled
  image, its weight image (``np.equal(wht, 0.0)`` becomes the DQ array) and every science
  extension of the input image, minus its sky (align.py:308-320, 464-483);
* ``fits.getdata(filename, ext=0)`` for SExtractor's segmentation image (catalogs.py:738-757);
* the SExtractor text catalogue and the selection rules of ``SExImageCatalog.execute``
  (catalogs.py:539-567, 823-898, 405-428, 968-988).

Here the FITS header is parsed on the host (FITS standard 4.0, section 4), the data block
travels AS IT IS IN THE FILE through two pinned staging buffers (the read of chunk k+1
overlaps the upload of chunk k) and ONE kernel (``spx_fits_decode``, ``csrc/spx_ingest.cu``)
turns the big-endian payload into the native array, applies BSCALE / BZERO the way astropy
does, subtracts the sky and writes the "weight == 0" plane - the pixels never exist on the
host in decoded form.  The returned CUDA tensors are what ``align.frame_displacements``,
``batch.gather_cutouts`` and ``batch.segment_boxes`` take.

There is no CPU decode path: without the CUDA library / a device ``read`` raises.
astropy is not available here, so the dtype / scaling rules below follow its documented
behaviour (``ImageHDU.data``: ``do_not_scale_image_data=False``, ``uint=True``) and the
FITS standard, and are tested against numpy's big-endian dtypes - parity with astropy
itself is unpinned.
"""
import os
import re
from collections import OrderedDict

import numpy as np
import torch

from . import _lib as L
from . import batch
from .utils import py2round

__all__ = ["parse_file_name", "FitsFile", "read_image", "load_drizzled", "load_image_ext",
           "load_segmentation", "get_ext_list", "read_sextractor_catalog", "select_sources",
           "DEFAULT_FILTERS"]

BLOCK = 2880
CARD = 80
_CHUNK = 32 << 20                     # bytes per pinned staging buffer


# --------------------------------------------------------------------------
# file names (utils.py:21-93)
# --------------------------------------------------------------------------
def parse_file_name(image_name):
    """``'j1234568q_flt.fits[sci,2]'`` -> ``('j1234568q_flt.fits', ('sci', 2))``;
    ``'x.fits[1]'`` -> ``('x.fits', 1)``; no brackets -> ``(name, None)``.
    Same results and ``ValueError`` messages as utils.py:50-93."""
    nbra, nket = image_name.count('['), image_name.count(']')
    if nbra + nket > 0:
        image_name = image_name.rstrip()
    if nbra != nket or nbra > 1 or (nbra == 1 and not image_name.endswith(']')):
        raise ValueError("Misplaced, unbalanced, or nested "
                         "brackets have been detected.")
    if nbra == 0:
        return image_name, None
    pos = image_name.find('[')
    if pos == 0:
        raise ValueError("No valid file name provided.")
    parts = image_name[pos + 1:-1].split(',')
    if len(parts) > 2:
        raise ValueError("Invalid extension specification.")
    try:
        number = int(parts[-1])
    except ValueError:
        raise ValueError("Invalid extension specification.")
    if len(parts) == 1:
        return image_name[:pos], number
    return image_name[:pos], (parts[0].strip(), number)


# --------------------------------------------------------------------------
# FITS headers (FITS standard 4.0, section 4: 80-byte cards, 2880-byte blocks)
# --------------------------------------------------------------------------
_NUM = re.compile(r'^[+-]?(\d+\.?\d*|\.\d+)([EeDd][+-]?\d+)?$')


def _card_value(text):
    """Value field of a card (bytes 11-80) -> Python value."""
    t = text.lstrip()
    if t.startswith("'"):                       # character string, '' = embedded quote
        out, i = [], 1
        while i < len(t):
            if t[i] == "'":
                if i + 1 < len(t) and t[i + 1] == "'":
                    out.append("'")
                    i += 2
                    continue
                break
            out.append(t[i])
            i += 1
        return "".join(out).rstrip()
    t = t.split('/', 1)[0].strip()
    if t == 'T':
        return True
    if t == 'F':
        return False
    if not t:
        return None
    if re.match(r'^[+-]?\d+$', t):
        return int(t)
    if _NUM.match(t):
        return float(t.replace('D', 'E').replace('d', 'e'))
    return t                                    # complex numbers etc.: left as text


def _parse_header(blocks):
    """Cards of one header (bytes, a multiple of 2880) -> OrderedDict (first occurrence
    of a keyword wins; commentary cards are skipped) and whether END was seen."""
    hdr, end = OrderedDict(), False
    for k in range(0, len(blocks), CARD):
        card = blocks[k:k + CARD].decode('ascii', 'replace')
        key = card[:8].rstrip()
        if key == 'END':
            end = True
            break
        if card[8:10] != '= ' or key in ('COMMENT', 'HISTORY', ''):
            continue
        if key not in hdr:
            hdr[key] = _card_value(card[10:])
    return hdr, end


class FitsHDU(object):
    """One header-data unit: header keywords, payload geometry, position in the file."""

    def __init__(self, index, header, data_offset):
        self.index = index
        self.header = header
        self.data_offset = data_offset
        self.is_image = index == 0 or str(header.get('XTENSION', '')).strip().upper() == 'IMAGE'
        self.name = str(header.get('EXTNAME', 'PRIMARY' if index == 0 else '')).strip().upper()
        self.ver = int(header.get('EXTVER', 1))
        self.bitpix = int(header['BITPIX'])
        naxis = int(header['NAXIS'])
        axes = [int(header['NAXIS%d' % (k + 1)]) for k in range(naxis)]
        self.shape = tuple(reversed(axes))                 # numpy order
        nelem = int(np.prod(axes, dtype=np.int64)) if naxis else 0
        if index > 0:
            nelem = int(header.get('GCOUNT', 1)) * (int(header.get('PCOUNT', 0)) + nelem)
        self.size = nelem if self.is_image else 0          # elements of the image
        self.data_bytes = nelem * abs(self.bitpix) // 8
        self.bscale = float(header.get('BSCALE', 1.0))
        self.bzero = float(header.get('BZERO', 0.0))

    @property
    def padded_bytes(self):
        return (self.data_bytes + BLOCK - 1) // BLOCK * BLOCK


_FLOAT_OUT = {8: torch.float32, 16: torch.float32, 32: torch.float64, 64: torch.float64,
              -32: torch.float32, -64: torch.float64}
_RAW_INT = {8: torch.uint8, 16: torch.int16, 32: torch.int32, 64: torch.int64}
_UNSIGNED = {16: getattr(torch, "uint16", None), 32: getattr(torch, "uint32", None),
             64: getattr(torch, "uint64", None)}
_MASK_KIND = {torch.uint8: L.T_I8, torch.int8: L.T_I8, torch.bool: L.T_I8,
              torch.int32: L.T_I32, torch.int64: L.T_I64}
_staging = {}


def _pinned_pair(device):
    """Two pinned staging buffers (each with the event of its last upload) and the upload
    stream of a device."""
    st = _staging.get(device.index)
    if st is None:
        st = {"bufs": [(torch.empty(_CHUNK, dtype=torch.uint8).pin_memory(), torch.cuda.Event())
                       for _ in range(2)],
              "stream": torch.cuda.Stream(device=device)}
        _staging[device.index] = st
    return st["bufs"], st["stream"]


READERS = max(2, min(8, (os.cpu_count() or 4) // 2))     # file-reader threads per staging chunk
_reader_pool = None


def _readers():
    global _reader_pool
    if _reader_pool is None:
        from concurrent.futures import ThreadPoolExecutor
        _reader_pool = ThreadPoolExecutor(max_workers=READERS, thread_name_prefix="spx-fits")
    return _reader_pool


def _pread_into(fd, view, offset, path):
    got = 0
    while got < len(view):
        r = os.preadv(fd, [view[got:]], offset + got)
        if not r:
            raise IOError("%s: unexpected end of file" % path)
        got += r


def release_staging():
    """Free the pinned staging buffers (2 x 32 MiB per device)."""
    _staging.clear()


def output_dtype(hdu, uint=True):
    """torch dtype of ``hdu.data`` under astropy's default rules: unscaled payloads keep
    their type; ``BSCALE = 1`` with ``BZERO = 2**(BITPIX-1)`` is the unsigned-integer
    convention (``uint=True``; BITPIX 8 with BZERO -128 is int8); any other scaling gives
    float32 for BITPIX 8 / 16 / -32 and float64 for 32 / 64 / -64."""
    bp = hdu.bitpix
    if hdu.bscale == 1.0 and hdu.bzero == 0.0:
        return _RAW_INT[bp] if bp > 0 else _FLOAT_OUT[bp]
    if uint and bp > 0 and hdu.bscale == 1.0:
        if bp == 8 and hdu.bzero == -128.0:
            return torch.int8
        if bp > 8 and hdu.bzero == float(1 << (bp - 1)) and _UNSIGNED[bp] is not None:
            return _UNSIGNED[bp]
    return _FLOAT_OUT[bp]


class FitsFile(object):
    """Header table of a FITS file + device reads of its image HDUs.

    >>> f = FitsFile('j9irw1rqq_flt.fits')
    >>> sci = f.read(('SCI', 1))                    # 2-D CUDA tensor, native byte order
    >>> f.header(0)['EXPTIME']
    """

    def __init__(self, path):
        self.path = path
        self.hdus = []
        with open(path, 'rb') as f:
            size = os.fstat(f.fileno()).st_size
            pos = 0
            while pos < size:
                chunks, end = [], False
                while not end:
                    blk = f.read(BLOCK)
                    if not chunks and not blk.strip(b'\0 '):
                        break                             # zero / blank padding after the last HDU
                    if len(blk) < BLOCK:
                        raise ValueError("%s: truncated FITS header at byte %d" % (path, pos))
                    hdr_part, end = _parse_header(blk)
                    chunks.append(blk)
                if not chunks:
                    break
                hdr, _ = _parse_header(b"".join(chunks))
                first = next(iter(hdr), None)
                if first != ('SIMPLE' if not self.hdus else 'XTENSION'):
                    if not self.hdus:
                        raise ValueError("%s is not a FITS file (no SIMPLE card)" % path)
                    break                                  # special records after the HDUs
                pos += len(chunks) * BLOCK
                hdu = FitsHDU(len(self.hdus), hdr, pos)
                if pos + hdu.data_bytes > size:
                    raise ValueError("%s: HDU %d is truncated" % (path, hdu.index))
                self.hdus.append(hdu)
                pos += hdu.padded_bytes
                f.seek(pos)
        if not self.hdus:
            raise ValueError("%s is not a FITS file" % path)

    def __len__(self):
        return len(self.hdus)

    def index_of(self, ext=None):
        """HDU number of an extension spec: an integer, ``'SCI'`` (first HDU of that name),
        ``('SCI', 2)`` (name + EXTVER, names compared case-insensitively), or ``None`` =
        the first HDU that holds data (what ``fits.getdata(filename)`` picks)."""
        if ext is None:
            for h in self.hdus:
                if h.is_image and h.size > 0:
                    return h.index
            raise IndexError("no image data in %s" % self.path)
        if isinstance(ext, (int, np.integer)) and not isinstance(ext, bool):
            k = int(ext) if ext >= 0 else len(self.hdus) + int(ext)
            if not 0 <= k < len(self.hdus):
                raise IndexError("%s has no HDU %d" % (self.path, ext))
            return k
        name, ver = (ext, None) if isinstance(ext, str) else (ext[0], int(ext[1]))
        name = name.strip().upper()
        for h in self.hdus:
            if h.name == name and (ver is None or h.ver == ver):
                return h.index
        raise KeyError("Extension %r not found in %s" % (ext, self.path))

    def hdu(self, ext=None):
        return self.hdus[self.index_of(ext)]

    def header(self, ext=0):
        return self.hdus[self.index_of(ext)].header

    def ext_list(self, extname='SCI'):
        """``[('SCI', 1), ('SCI', 2)]`` (utils.get_ext_list, utils.py:95-141)."""
        if not isinstance(extname, str):
            raise TypeError("Argument 'extname' must be either a string "
                            "indicating the value of the 'EXTNAME' keyword of the "
                            "extensions whose versions are to be returned.")
        extname = extname.upper()
        return [(extname, h.ver) for h in self.hdus
                if 'EXTNAME' in h.header and h.name == extname]

    # ---- payload ---------------------------------------------------------
    def read_raw(self, ext=None, device=None):
        """Upload the data block of an HDU unchanged: uint8 CUDA tensor of
        ``data_bytes`` rounded up to 16 (file -> pinned staging -> device, the file read
        of one chunk overlapping the upload of the previous one)."""
        batch.require_cuda()
        hdu = self.hdu(ext)
        device = torch.device('cuda', torch.cuda.current_device()) if device is None \
            else torch.device(device)
        nbytes = hdu.data_bytes
        raw = torch.empty((nbytes + 15) // 16 * 16, dtype=torch.uint8, device=device)
        if nbytes == 0:
            return raw
        bufs, side = _pinned_pair(device)
        cur = torch.cuda.current_stream(device)
        side.wait_stream(cur)                              # raw is allocated on `cur`
        fd = os.open(self.path, os.O_RDONLY)
        try:
            done, k = 0, 0
            while done < nbytes:
                pin, ev = bufs[k & 1]
                ev.synchronize()                           # its previous upload has left the buffer
                want = min(_CHUNK, nbytes - done)
                view = memoryview(pin.numpy())[:want]
                # the chunk is read by READERS threads at once (pread releases the GIL; one
                # thread copies out of the page cache at about 6 GB/s, a PCIe link takes 50)
                part = -(-want // READERS // 4096) * 4096
                jobs = [_readers().submit(_pread_into, fd, view[a:min(a + part, want)],
                                          hdu.data_offset + done + a, self.path)
                        for a in range(0, want, part)]
                for j in jobs:
                    j.result()
                with torch.cuda.stream(side):
                    raw[done:done + want].copy_(pin[:want], non_blocking=True)
                    ev.record(side)
                done += want
                k += 1
        finally:
            os.close(fd)
        cur.wait_stream(side)
        raw.record_stream(side)
        return raw

    def read(self, ext=None, device=None, sky=None, zero_mask=None, dtype=None, uint=True):
        """Image of an HDU as a CUDA tensor in native byte order.

        ``dtype``: torch dtype of the result (default: astropy's rule, see
        ``output_dtype``); ``sky``: a number subtracted from every pixel in the array's own
        precision (``img_sci -= image_sky[ext]``, align.py:480-481; floating results
        only); ``zero_mask``: a torch integer / bool dtype - also return a plane of that
        type that is 1 where the decoded pixel equals 0 (``np.equal(wht, 0.0)``,
        align.py:312).  Returns the tensor, or ``(tensor, mask)``."""
        hdu = self.hdu(ext)
        if not hdu.is_image:
            raise TypeError("HDU %d of %s is not an image" % (hdu.index, self.path))
        out_dtype = output_dtype(hdu, uint) if dtype is None else dtype
        raw = self.read_raw(hdu.index, device)
        n, bp = hdu.size, hdu.bitpix
        width = abs(bp) // 8
        floating = out_dtype in (torch.float32, torch.float64)
        bscale, bzero = hdu.bscale, hdu.bzero
        if floating:
            kind = L.T_F32 if out_dtype == torch.float32 else L.T_F64
            store = out_dtype
        else:
            if bp < 0 or torch.empty((), dtype=out_dtype).element_size() != width:
                raise TypeError("integer output needs an integer payload of the same width")
            if sky:
                raise TypeError("sky subtraction needs a floating-point result")
            if out_dtype == _RAW_INT[bp]:
                bscale, bzero = 1.0, 0.0               # do_not_scale_image_data
            elif not (bscale == 1.0 and bzero == float(int(bzero))):
                raise TypeError("BSCALE / BZERO of this HDU need a floating-point result")
            kind = {1: L.T_I8, 2: L.T_I16, 4: L.T_I32, 8: L.T_I64}[width]
            store = out_dtype
        esize = torch.empty((), dtype=store).element_size()
        if esize == width:                                 # decode in place
            out = raw
        else:
            out = torch.empty((n * esize + 15) // 16 * 16, dtype=torch.uint8, device=raw.device)
        mask = None
        if zero_mask is not None:
            if zero_mask not in _MASK_KIND:
                raise TypeError("zero_mask must be uint8, int8, bool, int32 or int64")
            msize = torch.empty((), dtype=zero_mask).element_size()
            mask = torch.empty((n * msize + 15) // 16 * 16, dtype=torch.uint8, device=raw.device)
        with torch.cuda.device(raw.device):
            L.check(L.load().spx_fits_decode(
                batch._ptr(raw), n, bp, bscale, bzero, float(sky) if sky else 0.0, kind,
                batch._ptr(out), _MASK_KIND.get(zero_mask, 0), batch._ptr(mask), batch._stream()))
        shape = hdu.shape if int(np.prod(hdu.shape, dtype=np.int64)) == n else (n,)
        res = out[:n * esize].view(store).view(shape)
        if mask is not None:
            return res, mask[:n * msize].view(zero_mask).view(shape)
        return res


def read_image(image_name, ext=None, **kwargs):
    """``fits.getdata(filename, ext)`` onto the device.  ``image_name`` may carry the
    extension in brackets (parse_file_name); ``ext`` overrides it."""
    fn, fext = parse_file_name(image_name)
    return FitsFile(fn).read(fext if ext is None else ext, **kwargs)


def get_ext_list(image, extname='SCI'):
    """utils.get_ext_list (utils.py:95-141) for a file name or an open ``FitsFile``."""
    f = image if isinstance(image, FitsFile) else FitsFile(image)
    return f.ext_list(extname)


def _units(bunit):
    return 'rate' if '/' in str(bunit) else 'counts'      # align.py:325, 473, 480


def load_drizzled(output_sci, output_wht=None, device=None, dq_dtype=torch.int64):
    """The reads of align.py:308-320 / 464-473: drizzled science image, exposure time
    (the extension's ``EXPTIME`` if present, else the primary's), units from ``BUNIT``,
    and - when a weight image is given - ``whtdq = (wht == 0)`` as an integer plane of
    ``dq_dtype`` (the reference's ``astype(np.int)`` is int64; uint8 is enough for the
    device pipeline).  Returns a dict: data, whtdq, exptime, units, header, file."""
    whtdq = None
    if output_wht is not None:
        fn, ext = parse_file_name(output_wht)
        wf = FitsFile(fn)
        # the decoded weights themselves are not needed: they are decoded in place
        _, whtdq = wf.read(ext, device=device, zero_mask=dq_dtype)
    fn, ext = parse_file_name(output_sci)
    f = FitsFile(fn)
    hdu = f.hdu(ext)
    hdr = hdu.header
    exptime = hdr['EXPTIME'] if 'EXPTIME' in hdr else f.header(0)['EXPTIME']
    return {"data": f.read(hdu.index, device=device), "whtdq": whtdq, "exptime": exptime,
            "units": _units(hdr['BUNIT']), "header": hdr, "file": f}


def load_image_ext(image, ext, sky=None, device=None):
    """One science extension of an input image as align.py:476-483 reads it: data minus
    ``sky`` (when not None), exposure time from the PRIMARY header, units from the
    extension's ``BUNIT``.  ``image``: file name or ``FitsFile``."""
    f = image if isinstance(image, FitsFile) else FitsFile(image)
    hdr = f.header(ext)
    data = f.read(ext, device=device, sky=sky if sky is not None else 0.0,
                  dtype=None if sky is None else _FLOAT_OUT[f.hdu(ext).bitpix])
    return {"data": data, "exptime": f.header(0)['EXPTIME'], "units": _units(hdr['BUNIT']),
            "header": hdr, "file": f}


def load_segmentation(filename, device=None):
    """SExtractor's segmentation check image, ``fits.getdata(filename, ext=0)``
    (catalogs.py:738-757), as an int32 / int64 CUDA tensor ready for
    ``batch.segment_boxes`` and the gather's segment test."""
    f = FitsFile(filename)
    hdu = f.hdu(0)
    seg = f.read(0, device=device, dtype=_RAW_INT[hdu.bitpix] if hdu.bitpix > 0 else None)
    if seg.dtype not in (torch.int32, torch.int64):
        seg = seg.to(torch.int32)
    return seg


# --------------------------------------------------------------------------
# SExtractor catalogue (catalogs.py:539-567) and source selection (catalogs.py:823-898)
# --------------------------------------------------------------------------
_FWHM2SIGMA = 2 * np.sqrt(2.0 * np.log(2.0))               # catalogs.py:31
_CATMAP = OrderedDict([                                     # catalogs.py:76-86
    ('NUMBER', 'id'), ('X_IMAGE', 'x'), ('Y_IMAGE', 'y'), ('FLUX_ISO', 'flux'),
    ('FLUXERR_ISO', 'fluxerr'), ('A_IMAGE', 'semi-major-a'), ('B_IMAGE', 'semi-major-b'),
    ('FWHM_IMAGE', 'fwhm'), ('CLASS_STAR', 'stellarity')])
REQUIRED_COLNAMES = ['id', 'x', 'y', 'flux', 'fluxerr', 'semi-major-a', 'semi-major-b',
                     'stellarity', 'fwhm']                  # catalogs.py:91-94
DEFAULT_FILTERS = [('flux', '>', 0), ('fwhm', '>', 0), ('semi-major-a', '>', 0),
                   ('semi-major-b', '>', 0), ('stellarity', '<=', 1.0)]   # catalogs.py:507-519


def read_sextractor_catalog(path, rename=True):
    """SExtractor ``ASCII_HEAD`` catalogue -> ``OrderedDict`` of numpy columns (int64 for
    columns whose every entry is an integer literal, else float64).  Header lines are
    ``#  <column number>  <NAME>  <description>  [unit]``; a gap in the numbering means the
    previous parameter is a vector and its extra columns are called ``NAME_1``, ``NAME_2``,
    ...  With ``rename`` the SExtractor names become the catalogue's own
    (``X_IMAGE`` -> ``x`` ..., catalogs.py:76-86, 563-567) and a missing required column
    raises the reference's ``ValueError`` (catalogs.py:554-558)."""
    declared, rows = [], []
    with open(path, 'r') as f:
        for line in f:
            s = line.strip()
            if not s:
                continue
            if s.startswith('#'):
                m = re.match(r'^#\s*(\d+)\s+([-\w]+)', s)
                if m:
                    declared.append((int(m.group(1)), m.group(2)))
                continue
            rows.append(s.split())
    ncol = len(rows[0]) if rows else (declared[-1][0] if declared else 0)
    names = []
    for k, (num, name) in enumerate(declared):
        nxt = declared[k + 1][0] if k + 1 < len(declared) else max(ncol, num) + 1
        names.append(name)
        names.extend("%s_%d" % (name, j) for j in range(1, nxt - num))
    names = names[:ncol] if rows else names
    if rows and any(len(r) != len(names) for r in rows):
        raise ValueError("%s: rows do not match the %d declared columns" % (path, len(names)))
    cat = OrderedDict()
    for j, name in enumerate(names):
        col = [r[j] for r in rows]
        if col and all(re.match(r'^[+-]?\d+$', v) for v in col):
            cat[name] = np.array([int(v) for v in col], dtype=np.int64)
        else:
            cat[name] = np.array([float(v) for v in col], dtype=np.float64)
    if rename:
        new_names = [_CATMAP.get(k, k) for k in cat]
        if not all(c in new_names for c in REQUIRED_COLNAMES):
            raise ValueError("Input raw catalog does not include all the "
                             "required columns.")
        cat = OrderedDict((_CATMAP.get(k, k), v) for k, v in cat.items())
    return cat


_CMP = {'>': np.greater, '>=': np.greater_equal, '==': np.equal, '!=': np.not_equal,
        '<': np.less, '<=': np.less_equal}                  # catalogs.py:62-71


def _select_rank(values, nrows, high):
    """catalogs.py:40-53: the ``nrows`` highest / lowest values (stable order; a negative
    count selects everything BUT that many from the other end)."""
    order = np.argsort(np.asarray(values), kind='stable')
    if high:
        idx = order[-nrows:] if nrows >= 0 else order[:-nrows]
    else:
        idx = order[:nrows] if nrows >= 0 else order[nrows:]
    mask = np.zeros(len(order), dtype=bool)
    mask[idx] = True
    return mask


def _rows(cat, mask):
    return OrderedDict((k, v[mask]) for k, v in cat.items())


def select_sources(raw, mask=None, mask_type=None, filters=DEFAULT_FILTERS):
    """``SExImageCatalog.execute`` after SExtractor has run (catalogs.py:823-898) on a raw
    catalogue with the catalogue's own column names (``read_sextractor_catalog``):

    1. positions become 0-based (``x -= 1``, ``y -= 1``) and are rounded half away from
       zero (``py2round``) for the mask look-up;
    2. ``mask``: ``mask_type='coords'`` - an ``[m, 2]`` integer array of (x, y) pixels to
       drop; ``'image'`` - a 2-D boolean image, True = bad (sources outside it are dropped);
    3. comparison filters ``(key, op, value)`` on existing columns; then ``pos_std =
       fwhm / (2 sqrt(2 ln 2) flux / fluxerr)`` (catalogs.py:405-428) and ``weight =
       1 / pos_std**2`` (catalogs.py:968-988) are added and the comparison filters run
       again (they may name the derived columns);
    4. rank filters ``(key, 'h' | 'l', nrows)`` last, each evaluated on the same table.

    Returns an ``OrderedDict`` of numpy columns."""
    cat = OrderedDict((k, np.array(v)) for k, v in raw.items())
    n = len(next(iter(cat.values()))) if cat else 0
    cat['x'] = cat['x'] - 1
    cat['y'] = cat['y'] - 1
    xi = py2round(np.asarray(cat['x'])).astype(int) if n else np.zeros(0, dtype=int)
    yi = py2round(np.asarray(cat['y'])).astype(int) if n else np.zeros(0, dtype=int)
    if mask is None:
        keep = np.ones(n, dtype=bool)
    elif mask_type == 'coords':
        bad = np.asarray(mask).reshape(-1, 2)
        keep = np.ones(n, dtype=bool)
        for lo in range(0, n, 4096):                       # [chunk, m] comparisons at a time
            sl = slice(lo, lo + 4096)
            keep[sl] = ~((xi[sl, None] == bad[None, :, 0]) &
                         (yi[sl, None] == bad[None, :, 1])).any(axis=1)
    elif mask_type == 'image':
        m = np.asarray(mask)
        ymax, xmax = m.shape
        inside = (xi >= 0) & (xi < xmax) & (yi >= 0) & (yi < ymax)
        keep = np.zeros(n, dtype=bool)
        keep[inside] = np.logical_not(m[yi[inside], xi[inside]])
    else:
        raise ValueError("Unexpected value of '_mask_type'. Contact "
                         "software developer.")

    def compare(table, sel):
        for key, op, val in (f[:3] for f in filters):
            op = ''.join(op.split())
            if op not in ('h', 'l') and key in table:
                sel = sel * _CMP[op](table[key], val)
        return sel

    keep = compare(cat, keep)
    cat = _rows(cat, keep)
    cat['pos_std'] = np.asarray(cat['fwhm'] / (_FWHM2SIGMA * cat['flux'] / cat['fluxerr']))
    cat['weight'] = 1.0 / cat['pos_std'] ** 2
    keep = compare(cat, np.ones(len(cat['x']), dtype=bool))
    cat = _rows(cat, keep)
    keep = np.ones(len(cat['x']), dtype=bool)
    for key, op, val in (f[:3] for f in filters):
        op = ''.join(op.split())
        if op in ('h', 'l') and key in cat:
            keep = keep * _select_rank(cat[key], val, op == 'h')
    return _rows(cat, keep)
