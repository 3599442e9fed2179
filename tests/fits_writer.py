"""Minimal FITS writer for the ingest tests (FITS standard 4.0: 80-byte cards in
2880-byte header blocks, big-endian data padded with zeros to a block boundary).
Written independently of ``subpixal_b200.ingest`` - it is the other side of the format.
Test infrastructure only."""
import numpy as np

BLOCK = 2880


def _card(key, value, comment=""):
    if isinstance(value, bool):
        v = "%20s" % ("T" if value else "F")
    elif isinstance(value, (int, np.integer)):
        v = "%20d" % value
    elif isinstance(value, (float, np.floating)):
        v = "%20s" % repr(float(value)).upper()
    else:
        v = "'%-8s'" % str(value).replace("'", "''")
        v = "%-20s" % v
    text = "%-8s= %s" % (key, v)
    if comment:
        text += " / " + comment
    return ("%-80s" % text)[:80].encode("ascii")


def _header(cards):
    raw = b"".join(cards) + b"%-80s" % b"END"
    return raw + b" " * (-len(raw) % BLOCK)


_BITPIX = {"u1": 8, "i2": 16, "i4": 32, "i8": 64, "f4": -32, "f8": -64}


def hdu_bytes(data, primary, extname=None, extver=None, extra=(), pad_cards=0):
    """One image HDU: ``data`` (numpy array or None) is stored big-endian in its own
    type; ``extra`` = additional (key, value[, comment]) cards; ``pad_cards`` COMMENT
    cards push the header over several blocks."""
    cards = [_card("SIMPLE", True) if primary else _card("XTENSION", "IMAGE")]
    if data is None:
        cards += [_card("BITPIX", 8), _card("NAXIS", 0)]
        payload = b""
    else:
        a = np.ascontiguousarray(data)
        code = a.dtype.kind.replace("b", "u") + str(a.dtype.itemsize)
        cards += [_card("BITPIX", _BITPIX[code]), _card("NAXIS", a.ndim)]
        cards += [_card("NAXIS%d" % (k + 1), n) for k, n in enumerate(reversed(a.shape))]
        payload = a.astype(a.dtype.newbyteorder(">")).tobytes()
    if primary:
        cards.append(_card("EXTEND", True))
    else:
        cards += [_card("PCOUNT", 0), _card("GCOUNT", 1)]
    if extname is not None:
        cards.append(_card("EXTNAME", extname, "extension name"))
    if extver is not None:
        cards.append(_card("EXTVER", extver))
    cards += [_card(*c) for c in extra]
    cards += [b"%-80s" % (b"COMMENT filler card %d" % k) for k in range(pad_cards)]
    return _header(cards) + payload + b"\0" * (-len(payload) % BLOCK)


def write_fits(path, hdus):
    """``hdus``: list of dicts with the keyword arguments of ``hdu_bytes`` (without
    ``primary``); the first one becomes the primary HDU."""
    with open(path, "wb") as f:
        for k, h in enumerate(hdus):
            f.write(hdu_bytes(primary=(k == 0), **h))
