"""Minimal FITS reader/writer exposing the small subset of the ``fitsio`` API that the
reference's I/O layer and CLI scripts use (test infrastructure only, never on the product path).

Supports exactly what occurs at the edges of the pair-counting path:
fixed-width BinTable columns of FITS types D/E/K/J/I/B/L/nD/nK/nA and fp64/fp32/int images,
optionally gzip-compressed.  No variable-length arrays, no scaling keywords.
"""
import gzip
import os

import numpy as np

BLOCK = 2880

_TFORM2DT = {"D": ">f8", "E": ">f4", "K": ">i8", "J": ">i4", "I": ">i2", "B": "u1", "L": "S1"}
_BITPIX2DT = {-64: ">f8", -32: ">f4", 64: ">i8", 32: ">i4", 16: ">i2", 8: "u1"}


def _parse_value(raw):
    raw = raw.strip()
    if raw.startswith("'"):
        end = 1
        out = []
        while end < len(raw):
            if raw[end] == "'":
                if end + 1 < len(raw) and raw[end + 1] == "'":
                    out.append("'")
                    end += 2
                    continue
                break
            out.append(raw[end])
            end += 1
        return "".join(out).rstrip()
    raw = raw.split("/")[0].strip()
    if raw == "T":
        return True
    if raw == "F":
        return False
    if raw == "":
        return None
    try:
        return int(raw)
    except ValueError:
        pass
    try:
        return float(raw.replace("D", "E"))
    except ValueError:
        return raw


class Header(dict):
    """dict with the couple of extra accessors fitsio headers offer."""

    def keys_list(self):
        return list(self.keys())


def _read_header(buf, pos):
    header = Header()
    while True:
        block = buf[pos:pos + BLOCK]
        if len(block) < BLOCK:
            raise EOFError
        pos += BLOCK
        done = False
        for k in range(0, BLOCK, 80):
            card = block[k:k + 80].decode("ascii", errors="replace")
            key = card[:8].strip()
            if key == "END":
                done = True
                break
            if card[8:10] == "= ":
                header[key] = _parse_value(card[10:])
            elif key == "HIERARCH" and "=" in card:
                # ESO HIERARCH convention for keywords longer than 8 characters (what fitsio
                # writes for WAVE_SOLUTION / DELTA_LAMBDA in the delta files)
                long_key, value = card[8:].split("=", 1)
                header[long_key.strip()] = _parse_value(value)
        if done:
            break
    return header, pos


def _tform_dtype(tform):
    tform = tform.strip()
    k = 0
    while k < len(tform) and tform[k].isdigit():
        k += 1
    rep = int(tform[:k]) if k else 1
    code = tform[k]
    if code == "A":
        return ("S%d" % rep), None
    base = _TFORM2DT[code]
    if rep == 1:
        return base, None
    return base, (rep,)


class _Column:
    def __init__(self, arr):
        self._arr = arr

    def __getitem__(self, item):
        return self._arr[item]

    def read(self):
        return self._arr


class HDU:
    def __init__(self, header, data_bytes):
        self._header = header
        self._raw = data_bytes
        self._table = None
        self._image = None

    # -- fitsio-like API
    def read_header(self):
        return self._header

    def get_extname(self):
        return self._header.get("EXTNAME", "")

    def _load(self):
        hdr = self._header
        if hdr.get("XTENSION", "").strip() == "BINTABLE":
            if self._table is None:
                nrows = hdr["NAXIS2"]
                fields = []
                for k in range(1, hdr["TFIELDS"] + 1):
                    base, shape = _tform_dtype(hdr["TFORM%d" % k])
                    name = hdr["TTYPE%d" % k].strip()
                    fields.append((name, base) if shape is None else (name, base, shape))
                dtype = np.dtype(fields)
                assert dtype.itemsize == hdr["NAXIS1"], (dtype.itemsize, hdr["NAXIS1"])
                self._table = np.frombuffer(self._raw, dtype=dtype, count=nrows)
            return self._table
        if self._image is None:
            naxis = hdr.get("NAXIS", 0)
            if naxis == 0:
                self._image = np.zeros(0)
            else:
                shape = tuple(hdr["NAXIS%d" % k] for k in range(naxis, 0, -1))
                count = int(np.prod(shape))
                self._image = np.frombuffer(self._raw, dtype=_BITPIX2DT[hdr["BITPIX"]],
                                            count=count).reshape(shape)
        return self._image

    def get_colnames(self):
        return list(self._load().dtype.names)

    def get_nrows(self):
        return int(self._header["NAXIS2"])

    def read(self, columns=None):
        arr = self._load()
        if arr.dtype.names is None:
            return arr.astype(arr.dtype.newbyteorder("="))
        native = np.dtype([(n, arr.dtype[n].newbyteorder("=") if arr.dtype[n].kind != "S"
                            else arr.dtype[n]) if arr.dtype[n].shape == () else
                           (n, arr.dtype[n].base.newbyteorder("="), arr.dtype[n].shape)
                           for n in arr.dtype.names])
        out = np.empty(arr.shape, dtype=native)
        for n in arr.dtype.names:
            out[n] = arr[n]
        return out

    def __getitem__(self, item):
        arr = self._load()
        if isinstance(item, str):
            col = arr[item]
            if col.dtype.kind != "S":
                col = col.astype(col.dtype.newbyteorder("="))
            return _Column(col)
        # image slicing
        return arr[item].astype(arr.dtype.newbyteorder("="))


class FITS:
    """``fitsio.FITS`` look-alike (read mode, and 'rw' + clobber write mode)."""

    def __init__(self, filename, mode="r", clobber=False):
        self._filename = os.path.expandvars(filename)
        self._mode = mode
        self._hdus = []
        self._pending = []
        if mode == "r":
            opener = gzip.open if self._filename.endswith(".gz") else open
            try:
                with opener(self._filename, "rb") as fin:
                    buf = fin.read()
            except FileNotFoundError as err:
                raise OSError(str(err)) from err
            pos = 0
            while pos < len(buf):
                try:
                    header, pos = _read_header(buf, pos)
                except EOFError:
                    break
                naxis = header.get("NAXIS", 0)
                size = 0
                if naxis > 0:
                    size = abs(header["BITPIX"]) // 8
                    for k in range(1, naxis + 1):
                        size *= header["NAXIS%d" % k]
                    size += header.get("PCOUNT", 0)
                self._hdus.append(HDU(header, buf[pos:pos + size]))
                pos += ((size + BLOCK - 1) // BLOCK) * BLOCK
        elif mode == "rw":
            if os.path.exists(self._filename) and not clobber:
                raise OSError("file exists")
        else:
            raise ValueError(mode)

    # -- container protocol
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False

    def __len__(self):
        return len(self._hdus)

    def __iter__(self):
        return iter(self._hdus)

    def __contains__(self, name):
        if isinstance(name, int):
            return 0 <= name < len(self._hdus)
        return any(h.get_extname().strip() == name for h in self._hdus)

    def __getitem__(self, item):
        if isinstance(item, slice):
            return self._hdus[item]
        if isinstance(item, (int, np.integer)):
            return self._hdus[item]
        for hdu in self._hdus:
            if hdu.get_extname().strip() == item:
                return hdu
        raise KeyError(item)

    # -- writing
    def write(self, data, names=None, comment=None, units=None, header=None, extname=None):
        # fitsio writes through immediately and the reference scripts do not always close()
        self._pending.append((data, names, units, header, extname))
        self._flush()

    def close(self):
        if self._mode == "rw":
            self._flush()
            self._mode = "closed"

    def _flush(self):
        out = bytearray()
        out += _cards_to_bytes([("SIMPLE", True), ("BITPIX", 8), ("NAXIS", 0), ("EXTEND", True)])
        for data, names, units, header, extname in self._pending:
            out += _table_bytes(data, names, units, header, extname)
        opener = gzip.open if self._filename.endswith(".gz") else open
        with opener(self._filename, "wb") as fout:
            fout.write(bytes(out))


def _fmt_card(key, value):
    if isinstance(value, (bool, np.bool_)):
        val = "%20s" % ("T" if value else "F")
    elif isinstance(value, (int, np.integer)):
        val = "%20d" % int(value)
    elif isinstance(value, (float, np.floating)):
        val = "%20s" % repr(float(value)).upper()
    elif value is None:
        val = " " * 20
    else:
        val = "'%-8s'" % str(value).replace("'", "''")
        val = "%-20s" % val
    card = "%-8s= %s" % (key[:8], val)
    return card[:80].ljust(80)


def _cards_to_bytes(cards):
    text = "".join(_fmt_card(k, v) for k, v in cards) + "END".ljust(80)
    pad = (-len(text)) % BLOCK
    return (text + " " * pad).encode("ascii")


def _table_bytes(data, names, units, header, extname):
    cols = [np.asarray(col) for col in data]
    nrows = cols[0].shape[0] if cols[0].ndim else 1
    fields, tforms = [], []
    for name, col in zip(names, cols):
        shape = col.shape[1:]
        rep = int(np.prod(shape)) if shape else 1
        if col.dtype.kind == "f":
            base, code = ">f8", "D"
        elif col.dtype.kind in "iu":
            base, code = ">i8", "K"
        elif col.dtype.kind == "b":
            base, code = "S1", "L"
        elif col.dtype.kind in "SU":  # fixed-width strings: nA (picca_metal_dmat.py's ABS_IGM)
            col = col.astype("S")
            cols[len(fields)] = col
            fields.append((name, col.dtype))
            tforms.append("%dA" % col.dtype.itemsize)
            continue
        else:
            raise TypeError(col.dtype)
        fields.append((name, base, shape) if shape else (name, base))
        tforms.append("%d%s" % (rep, code) if rep != 1 else code)
    dtype = np.dtype(fields)
    table = np.zeros(nrows, dtype=dtype)
    for name, col in zip(names, cols):
        table[name] = col
    cards = [("XTENSION", "BINTABLE"), ("BITPIX", 8), ("NAXIS", 2), ("NAXIS1", dtype.itemsize),
             ("NAXIS2", nrows), ("PCOUNT", 0), ("GCOUNT", 1), ("TFIELDS", len(cols))]
    for k, (name, tform) in enumerate(zip(names, tforms), start=1):
        cards.append(("TTYPE%d" % k, name))
        cards.append(("TFORM%d" % k, tform))
        if units is not None and units[k - 1]:
            cards.append(("TUNIT%d" % k, units[k - 1]))
    if extname is not None:
        cards.append(("EXTNAME", extname))
    for item in (header or []):
        cards.append((item["name"], item["value"]))
    raw = table.tobytes()
    pad = (-len(raw)) % BLOCK
    return _cards_to_bytes(cards) + raw + b"\0" * pad


def read(filename, ext=1, columns=None):
    """``fitsio.read`` look-alike: return HDU ``ext`` as a native-endian array."""
    with FITS(filename) as hdul:
        return hdul[ext].read()
