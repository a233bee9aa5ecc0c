"""Minimal FITS primary-HDU reader / writer (FITS standard 4.0, sections 3-4): enough for the on-disk contract of
MUFASA's fit products -- ``{root}_{n}vcomp.fits`` = concatenate([parcube, errcube]) with PLANEi cards
(mufasa/multi_v_fit.py:481-582, master_fitter.py:1700-1718).  astropy is not installed here; files
written by this module are plain single-HDU images any FITS reader opens, and files written by astropy / the
reference with a single image HDU are read back.  No WCS logic: header cards are carried as (value, comment) pairs.
"""
import numpy as np

BLOCK = 2880
_BITPIX = {-64: '>f8', -32: '>f4', 8: 'u1', 16: '>i2', 32: '>i4', 64: '>i8'}
_DTYPE_BITPIX = {'f8': -64, 'f4': -32, 'u1': 8, 'i2': 16, 'i4': 32, 'i8': 64}


class Header(dict):
    """Ordered keyword -> (value, comment); HISTORY / COMMENT lines are kept in lists."""

    def __init__(self):
        super().__init__()
        self.history, self.comments = [], []

    def set(self, keyword, value, comment=None):
        self[keyword.upper()] = (value, comment)

    def value(self, keyword, default=None):
        return self[keyword][0] if keyword in self else default


def _fmt_value(v):
    if isinstance(v, (bool, np.bool_)):
        return "%20s" % ("T" if v else "F")
    if isinstance(v, (int, np.integer)):
        return "%20d" % int(v)
    if isinstance(v, (float, np.floating)):
        s = repr(float(v)).upper()
        if "E" not in s and "." not in s and "N" not in s:
            s += "."
        return "%20s" % s
    s = str(v).replace("'", "''")
    return "'%-8s'" % s


def _card(key, value, comment):
    body = "%-8s= %s" % (key[:8], _fmt_value(value))
    if comment:
        body += " / " + str(comment)
    return body[:80].ljust(80)


def _text_cards(key, lines):
    out = []
    for ln in lines:
        ln = str(ln)
        for i in range(0, max(len(ln), 1), 72):
            out.append(("%-8s%s" % (key, ln[i:i + 72])).ljust(80))
    return out


def write(path, data, header=None):
    """Write ``data`` (ndarray, any of u1/i2/i4/i8/f4/f8) as the primary HDU.  ``header``: Header or dict of
    keyword -> value | (value, comment); structural keywords are regenerated from the array."""
    data = np.asarray(data)
    code = data.dtype.str[1:]
    if code not in _DTYPE_BITPIX:
        raise TypeError("unsupported dtype for FITS: %s" % data.dtype)
    bitpix = _DTYPE_BITPIX[code]
    cards = [_card("SIMPLE", True, "conforms to FITS standard"), _card("BITPIX", bitpix, "array data type"),
             _card("NAXIS", data.ndim, "number of array dimensions")]
    for i, n in enumerate(data.shape[::-1]):
        cards.append(_card("NAXIS%d" % (i + 1), n, None))
    skip = {"SIMPLE", "BITPIX", "NAXIS", "EXTEND", "END"} | {"NAXIS%d" % i for i in range(1, 10)}
    if header is not None:
        for k, v in header.items():
            if k.upper() in skip:
                continue
            val, com = v if isinstance(v, tuple) else (v, None)
            cards.append(_card(k.upper(), val, com))
        cards += _text_cards("HISTORY", getattr(header, "history", []))
        cards += _text_cards("COMMENT", getattr(header, "comments", []))
    cards.append("END".ljust(80))
    head = "".join(cards)
    head += " " * (-len(head) % BLOCK)
    raw = np.ascontiguousarray(data, dtype=_BITPIX[bitpix])   # the one copy: byte order (and layout, if strided)
    with open(path, "wb") as f:
        f.write(head.encode("ascii"))
        f.write(raw.reshape(-1).data)                          # straight from the array's buffer
        f.write(b"\0" * (-raw.nbytes % BLOCK))


def _parse_value(s):
    s = s.strip()
    if s.startswith("'"):
        end = 1
        while True:                      # closing quote ('' is an escaped quote)
            end = s.index("'", end)
            if s[end:end + 2] == "''":
                end += 2
                continue
            break
        return s[1:end].replace("''", "'").rstrip(), s[end + 1:]
    val, _, rest = s.partition("/")
    val = val.strip()
    rest = "/" + rest if rest else ""
    if val == "T":
        return True, rest
    if val == "F":
        return False, rest
    try:
        return int(val), rest
    except ValueError:
        try:
            return float(val.replace("D", "E")), rest
        except ValueError:
            return val, rest


def read(path):
    """Returns (data ndarray in native byte order, Header) of the primary HDU."""
    hdr = Header()
    with open(path, "rb") as f:
        done = False
        while not done:
            block = f.read(BLOCK)
            if len(block) < BLOCK:
                raise ValueError("truncated FITS header")
            for i in range(0, BLOCK, 80):
                card = block[i:i + 80].decode("ascii", errors="replace")
                key = card[:8].strip()
                if key == "END":
                    done = True
                    break
                if key == "HISTORY":
                    hdr.history.append(card[8:].rstrip())
                elif key == "COMMENT":
                    hdr.comments.append(card[8:].rstrip())
                elif card[8:10] == "= ":
                    val, rest = _parse_value(card[10:])
                    com = rest.partition("/")[2].strip() or None
                    hdr[key] = (val, com)
        naxis = int(hdr.value("NAXIS", 0))
        shape = tuple(int(hdr.value("NAXIS%d" % i)) for i in range(naxis, 0, -1))
        dt = np.dtype(_BITPIX[int(hdr.value("BITPIX"))])
        n = int(np.prod(shape)) if shape else 0
        data = np.frombuffer(f.read(n * dt.itemsize), dtype=dt, count=n).reshape(shape)
    bscale, bzero = hdr.value("BSCALE", 1), hdr.value("BZERO", 0)
    data = data.astype(dt.newbyteorder("="))
    if bscale != 1 or bzero != 0:
        data = data * bscale + bzero
    return data, hdr
