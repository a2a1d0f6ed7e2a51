"""B200 delta loader: ``read_deltas`` with the reference's signature and return tuple
(py/picca/io.py:383-512), reading the BinTable delta files straight into the SoA CSR buffers of
the pair kernels.

What the reference does per forest in Python -- one fitsio HDU object, four column reads, one
``Delta`` constructor, ``10**log_lambda``, two scipy interpolations, the weight evolution and
``project()`` (io.py:354-360, :493-507; data.py:375-474, :622-655) -- happens here per FILE:

  host   gunzip + ``pb2_fits_scan`` / ``pb2_fits_cards`` (plain C): HDU table and header cards;
  H2D    the raw big-endian file bytes, as they are;
  device ``pb2_delta_unpack`` (byte swap + de-interleave into log_lambda / delta / weights at the
         CSR offsets) and ``pb2_delta_prepare`` (z, r_comov, dist_m, weight evolution, projection).

The returned ``data`` is the reference's ``dict[healpix] -> list[Delta]``; every array attribute of
a ``Delta`` is a view into one contiguous host array per field.  There is no CPU fallback: without
the CUDA library or a device the call raises.

Parity: the table interpolation follows scipy's ``interp1d`` operation by operation, so r_comov
and dist_m are bit-equal to the reference's GIVEN the same z; ``10**x`` on the device (``exp10``)
can differ from NumPy's power in the last ulp, so by default z, r_comov and dist_m agree with the
reference within 4 ulp (16 ulp for files that store LAMBDA, whose log10 is taken on the device
too).  ``PICCA_B200_HOST_POW=1`` (parity mode, used by the golden tests) evaluates ``10**log_lambda / lambda_abs - 1`` with NumPy on the host instead; then z, r_comov and
dist_m are bit-equal.  Weights and projected deltas are within 1e-13 relative either way (``pow``
and re-associated sums).  Both on-disk flavours are read: one BinTable HDU per forest
(``Delta.from_fitsio``) and the ImageHDU layout (``Delta.from_image``, data.py:519-620: common
wavelength grid, METADATA table, 2-D images; ``pb2_delta_image_count`` / ``_unpack`` keep the
pixels with WEIGHT > 0).  ``rebin_factor`` (``Delta.rebin``, data.py:657-686) runs on the device
too (``pb2_delta_rebin``: bin sums in np.bincount's order, bit-equal given the same wavelengths;
the wavelengths ``10**log_lambda`` of that step are always NumPy's, because the reference's bin
count is decided by their last ulp).
"""
import ctypes
import glob
import gzip
import os
import sys
from concurrent.futures import ThreadPoolExecutor
from configparser import ConfigParser

import numpy as np

from . import _lib
from .engine import get_engine
from .forest import PIXEL_FIELDS, Delta, register_soa
from .synth import ang2pix_ring

_KEYS = (["EXTNAME", "RA", "DEC", "Z", "THING_ID", "PLATE", "MJD", "FIBERID", "LOS_ID",
          "BLINDING", "XTENSION"] + ["TTYPE%d" % k for k in range(1, 9)] +
         ["TFORM%d" % k for k in range(1, 9)])
_KEY_BYTES = "".join(k.ljust(8) for k in _KEYS).encode("ascii")
_K = {k: i for i, k in enumerate(_KEYS)}
_TFORM_BYTES = {"D": 8, "E": 4, "K": 8, "J": 4, "I": 2, "B": 1, "L": 1, "A": 1}
_UPLOAD_BATCH = 256 << 20  # raw bytes staged per H2D copy


def userprint(*args, **kwds):
    """reference py/picca/utils.py:31-40"""
    print(*args, **kwds)
    sys.stdout.flush()


# ------------------------------------------------------------------------------------ FITS (host)
def _file_bytes(path):
    with open(path, "rb") as f:
        raw = f.read()
    if raw[:2] == b"\x1f\x8b":
        raw = gzip.decompress(raw)
    return np.frombuffer(raw, dtype=np.uint8)


def _scan(buf):
    """HDU table of a FITS buffer: int64 [n_hdu, 8] (see pb2_fits_scan)."""
    lib = _lib.lib()
    cap = max(16, buf.size // 2880 + 1)
    info = np.zeros((cap, 8), dtype=np.int64)
    n = lib.pb2_fits_scan(buf.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(buf.size),
                          ctypes.c_int64(cap), info.ctypes.data_as(ctypes.c_void_p))
    if n < 0:
        raise OSError("picca_b200: " + lib.pb2_last_error().decode("utf-8", "replace"))
    return info[:n]


def _cards(buf, header_off, keys=None):
    """Header cards ``keys`` (default _KEYS) of the HDUs starting at ``header_off``:
    (kind, num, inum, strings)."""
    lib = _lib.lib()
    key_bytes = _KEY_BYTES if keys is None else "".join(k.ljust(8) for k in keys).encode("ascii")
    n, nk = len(header_off), len(_KEYS if keys is None else keys)
    kind = np.zeros((n, nk), dtype=np.int32)
    num = np.zeros((n, nk), dtype=np.float64)
    inum = np.zeros((n, nk), dtype=np.int64)
    strs = np.zeros((n, nk), dtype="S24")
    header_off = np.ascontiguousarray(header_off, dtype=np.int64)
    _lib.check(lib.pb2_fits_cards(
        buf.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(buf.size), ctypes.c_int64(n),
        header_off.ctypes.data_as(ctypes.c_void_p), ctypes.c_int32(nk),
        ctypes.c_char_p(key_bytes), kind.ctypes.data_as(ctypes.c_void_p),
        num.ctypes.data_as(ctypes.c_void_p), inum.ctypes.data_as(ctypes.c_void_p),
        strs.ctypes.data_as(ctypes.c_void_p)), "pb2_fits_cards")
    return kind, num, inum, strs


def _column_offsets(ttypes, tforms):
    """name -> (byte offset, type letter, repeat) of a BinTable row."""
    out, pos = {}, 0
    for name, form in zip(ttypes, tforms):
        if not form:
            break
        form = form.strip()
        digits = "".join(c for c in form if c.isdigit())
        letter = form[len(digits):len(digits) + 1]
        if letter not in _TFORM_BYTES:
            raise NotImplementedError("picca_b200.io: TFORM %r is not supported" % form)
        rep = int(digits) if digits else 1
        out[name.strip().upper()] = (pos, letter, rep)
        pos += rep * _TFORM_BYTES[letter]
    return out, pos


def _hierarch(buf, header_off, key):
    """Value of a long (HIERARCH) keyword of one header, or KeyError like fitsio's header."""
    lib = _lib.lib()
    kind, num = ctypes.c_int32(0), ctypes.c_double(0.)
    text = ctypes.create_string_buffer(24)
    _lib.check(lib.pb2_fits_hierarch(
        buf.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(buf.size), ctypes.c_int64(header_off),
        ctypes.c_char_p(key.encode("ascii")), ctypes.byref(kind), ctypes.byref(num), text),
        "pb2_fits_hierarch")
    if kind.value == 0:
        raise KeyError(key)
    return num.value if kind.value == 1 else text.value.decode("ascii", "replace")


def _open_delta_file(path, z_min_qso, z_max_qso, rebin=False):
    """One delta file -> _FileForests (BinTable flavour) or _ImageForests (ImageHDU flavour);
    io.py:354-360: an extension called LAMBDA selects Delta.from_image.  With ``rebin`` the
    wavelength solution of the file is read like io.py:362-373 (header of the LAMBDA extension,
    else of HDU 1)."""
    buf = _file_bytes(path)
    info = _scan(buf)
    kind, num, inum, strs = _cards(buf, info[:, 0])
    names = [s.decode("ascii", "replace") for s in strs[:, _K["EXTNAME"]]]
    image = "LAMBDA" in names
    dwave = None
    if rebin:
        card = names.index("LAMBDA") if image else 1
        if _hierarch(buf, int(info[card, 0]), "WAVE_SOLUTION") != 'lin':
            raise ValueError('Delta rebinning only implemented for linear lambda bins')
        dwave = float(_hierarch(buf, int(info[card, 0]), "DELTA_LAMBDA"))
    if image:
        part = _ImageForests(path, buf, info, names, z_min_qso, z_max_qso)
    else:
        part = _FileForests(path, buf, info, kind, num, inum, strs, z_min_qso, z_max_qso)
    part.dwave = dwave
    return part


class _ImageForests:
    """ImageHDU flavour (Delta.from_image, data.py:519-620): a common wavelength grid, a METADATA
    table and 2-D DELTA / WEIGHT images; a forest keeps its pixels with WEIGHT > 0."""
    is_image = True

    def __init__(self, path, buf, info, names, z_min_qso, z_max_qso):
        self.path, self.buf = path, buf
        hdu = {name: k for k, name in reversed(list(enumerate(names))) if name}
        if "METADATA" not in hdu:
            raise KeyError("METADATA")
        meta = hdu["METADATA"]
        n_col = int(info[meta, 7])
        keys = ["BLINDING"] + ["TTYPE%d" % k for k in range(1, n_col + 1)] + \
            ["TFORM%d" % k for k in range(1, n_col + 1)]
        _, _, _, strs = _cards(buf, info[meta:meta + 1, 0], keys)
        dec = [x.decode("ascii", "replace") for x in strs[0]]
        blinding = dec[0] if dec[0] else "none"                       # data.py:546-554
        delta_name = "DELTA" if blinding == "none" else "DELTA_BLIND"
        cols, width = _column_offsets(dec[1:1 + n_col], dec[1 + n_col:1 + 2 * n_col])
        if width != info[meta, 5]:
            raise OSError("picca_b200.io: METADATA row width mismatch in %s" % path)
        n_forest = int(info[meta, 6])
        table = np.frombuffer(buf, dtype=np.uint8, count=n_forest * width,
                              offset=int(info[meta, 1])).reshape(n_forest, width)

        def column(name):
            off, letter, rep = cols[name]
            dt = {"D": ">f8", "E": ">f4", "K": ">i8", "J": ">i4", "I": ">i2", "B": "u1"}[letter]
            nbytes = np.dtype(dt).itemsize
            return np.ascontiguousarray(table[:, off:off + nbytes]).view(dt).reshape(n_forest)

        if "LOGLAM" in hdu:                                            # data.py:558-563
            wave, self.wave_flag = hdu["LOGLAM"], False
        else:
            wave, self.wave_flag = hdu["LAMBDA"], True
        for name in (delta_name, "WEIGHT"):
            if name not in hdu:
                raise KeyError(name)
        d, w = hdu[delta_name], hdu["WEIGHT"]
        self.n_lambda = int(info[wave, 5])
        for k in (d, w):
            if info[k, 3] != -64 or info[k, 5] != self.n_lambda or info[k, 6] != n_forest:
                raise NotImplementedError("picca_b200.io: image HDUs must be fp64 "
                                          "[n_forest][n_lambda]; file %s" % path)
        if info[wave, 3] != -64:
            raise NotImplementedError("picca_b200.io: the wavelength grid must be fp64")
        self.lambda_off, self.delta_off, self.weight_off = (int(info[k, 1]) for k in (wave, d, w))
        if "THING_ID" in cols:                                         # data.py:575-586
            los_id, plate, mjd, fiberid = (column(c) for c in ("THING_ID", "PLATE", "MJD",
                                                               "FIBERID"))
        elif "LOS_ID" in cols:
            los_id = plate = mjd = fiberid = column("LOS_ID")
        else:
            raise Exception("Could not find THING_ID or LOS_ID")
        z = column("Z").astype(np.float64)
        keep = np.nonzero((z >= z_min_qso) & (z <= z_max_qso))[0]     # data.py:602, inclusive
        self.rows = keep.astype(np.int32)
        self.n = len(keep)
        self.ra = column("RA").astype(np.float64)[keep]
        self.dec = column("DEC").astype(np.float64)[keep]
        self.z_qso = z[keep]
        self.los_id, self.plate = los_id.astype(np.int64)[keep], plate.astype(np.int64)[keep]
        self.mjd, self.fiberid = mjd.astype(np.int64)[keep], fiberid.astype(np.int64)[keep]
        self.wave_is_lambda = np.full(self.n, self.wave_flag, dtype=bool)
        self.n_pix = None  # known after the device count (count_pixels)
        self.d_raw = None

    def count_pixels(self, eng):
        """upload the file and count the WEIGHT > 0 pixels of every kept forest"""
        torch = eng.torch
        self.d_raw = torch.from_numpy(np.array(self.buf, copy=True)).to(eng.device)
        self.d_rows = torch.from_numpy(self.rows).to(eng.device)
        count = torch.zeros(max(self.n, 1), dtype=torch.int32, device=eng.device)
        _lib.check(eng.lib.pb2_delta_image_count(
            ctypes.c_int64(self.n), ctypes.c_void_p(self.d_raw.data_ptr()),
            ctypes.c_int64(self.weight_off), ctypes.c_int32(self.n_lambda),
            ctypes.c_void_p(self.d_rows.data_ptr()), ctypes.c_void_p(count.data_ptr()),
            eng.stream_ptr()), "pb2_delta_image_count")
        self.n_pix = count[:self.n].cpu().numpy().astype(np.int64)
        self.buf = None

    def unpack(self, eng, n_take, d_offset_ptr, d_ll, d_delta, d_w):
        _lib.check(eng.lib.pb2_delta_image_unpack(
            ctypes.c_int64(n_take), ctypes.c_void_p(self.d_raw.data_ptr()),
            ctypes.c_int64(self.lambda_off), ctypes.c_int64(self.delta_off),
            ctypes.c_int64(self.weight_off), ctypes.c_int32(self.n_lambda),
            ctypes.c_void_p(self.d_rows.data_ptr()), ctypes.c_void_p(d_offset_ptr),
            ctypes.c_void_p(d_ll.data_ptr()), ctypes.c_void_p(d_delta.data_ptr()),
            ctypes.c_void_p(d_w.data_ptr()), eng.stream_ptr()), "pb2_delta_image_unpack")
        eng.torch.cuda.current_stream().synchronize()
        self.d_raw = self.d_rows = None


class _FileForests:
    """BinTable flavour: per kept forest the row geometry and the header values."""
    is_image = False

    def __init__(self, path, buf, info, kind, num, inum, strs, z_min_qso, z_max_qso):
        self.path = path
        hdus = np.arange(1, len(info))  # hdul[1:]
        z = num[hdus, _K["Z"]]
        if np.any(kind[hdus, _K["Z"]] != 1):
            raise KeyError("Z")
        keep = hdus[(z_min_qso < z) & (z < z_max_qso)]  # io.py:359-360, strict
        self.buf = buf
        self.n = len(keep)
        self.row0 = info[keep, 1]
        self.row_bytes = info[keep, 5].astype(np.int32)
        self.n_pix = info[keep, 6]
        self.ra = num[keep, _K["RA"]]
        self.dec = num[keep, _K["DEC"]]
        self.z_qso = num[keep, _K["Z"]]
        if np.any(kind[keep, _K["RA"]] != 1) or np.any(kind[keep, _K["DEC"]] != 1):
            raise KeyError("RA")
        has_thing = kind[keep, _K["THING_ID"]] == 1
        has_los = kind[keep, _K["LOS_ID"]] == 1
        if np.any(~has_thing & ~has_los):
            raise Exception("Could not find THING_ID or LOS_ID")  # data.py:462-463
        pick = lambda key: np.where(has_thing, inum[keep, _K[key]], inum[keep, _K["LOS_ID"]])
        self.los_id = pick("THING_ID")
        self.plate, self.mjd, self.fiberid = pick("PLATE"), pick("MJD"), pick("FIBERID")
        # column layout: usually one per file; computed once per distinct (TTYPE, TFORM) set
        t0, f0 = _K["TTYPE1"], _K["TFORM1"]
        layout = np.concatenate([strs[keep, t0:t0 + 8], strs[keep, f0:f0 + 8],
                                 strs[keep, _K["BLINDING"]][:, None]], axis=1)
        self.col_off = np.zeros((self.n, 3), dtype=np.int32)
        self.wave_is_lambda = np.zeros(self.n, dtype=bool)
        uniq, inverse = np.unique(layout, axis=0, return_inverse=True) if self.n else ([], [])
        inverse = np.asarray(inverse).reshape(-1)
        for u, row in enumerate(uniq):
            dec = [s.decode("ascii", "replace") for s in row]
            cols, width = _column_offsets(dec[:8], dec[8:16])
            blinding = dec[16] if dec[16] else "none"            # data.py:395-400
            delta_name = "DELTA" if blinding == "none" else "DELTA_BLIND"
            if delta_name not in cols:
                raise KeyError(delta_name)
            if "LOGLAM" in cols:                                   # data.py:409-414
                wave, is_lambda = "LOGLAM", False
            elif "LAMBDA" in cols:
                wave, is_lambda = "LAMBDA", True
            else:
                raise KeyError("Did not find LOGLAM or LAMBDA in delta file")
            if "WEIGHT" not in cols:
                raise KeyError("WEIGHT")
            for name in (wave, delta_name, "WEIGHT"):
                if cols[name][1] != "D" or cols[name][2] != 1:
                    raise NotImplementedError("picca_b200.io: column %s is not a scalar fp64 "
                                              "column" % name)
            sel = inverse == u
            if np.any(self.row_bytes[sel] != width):
                raise OSError("picca_b200.io: NAXIS1 does not match the TFORM widths in %s" % path)
            self.col_off[sel] = (cols[wave][0], cols[delta_name][0], cols["WEIGHT"][0])
            self.wave_is_lambda[sel] = is_lambda


def find_order(in_dir, delta_attributes):
    """Order of the continuum polynomial from the delta-attributes file (io.py:31-112): header card
    FITORDER of HDU FIT_METADATA, else of STACK_DELTAS, else ``[expected flux] order`` of
    ``in_dir/../.config.ini``, else None."""
    if delta_attributes is None:
        delta_attributes = in_dir + "/../Log/delta_attributes.fits.gz"
        userprint(f"WARNING: delta_attributes file not given, setting to {delta_attributes}")
    userprint(f"Reading delta attributes from {delta_attributes}")

    def from_config():
        config = ConfigParser()
        config.read(in_dir + "/../.config.ini")
        if "expected flux" in config and "order" in config["expected flux"]:
            return config["expected flux"].getint("order")
        userprint("WARNING: `order` not found in delta config file")
        return None

    try:
        buf = _file_bytes(delta_attributes)
    except OSError as e:
        userprint(f"WARNING: OSError encountered: {str(e)}")
        order = from_config()
    else:
        info = _scan(buf)
        lib = _lib.lib()
        keys = "".join(k.ljust(8) for k in ("EXTNAME", "FITORDER")).encode("ascii")
        n = len(info)
        kind = np.zeros((n, 2), dtype=np.int32)
        num = np.zeros((n, 2))
        inum = np.zeros((n, 2), dtype=np.int64)
        strs = np.zeros((n, 2), dtype="S24")
        off = np.ascontiguousarray(info[:, 0])
        _lib.check(lib.pb2_fits_cards(
            buf.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(buf.size), ctypes.c_int64(n),
            off.ctypes.data_as(ctypes.c_void_p), ctypes.c_int32(2), ctypes.c_char_p(keys),
            kind.ctypes.data_as(ctypes.c_void_p), num.ctypes.data_as(ctypes.c_void_p),
            inum.ctypes.data_as(ctypes.c_void_p), strs.ctypes.data_as(ctypes.c_void_p)),
            "pb2_fits_cards")
        order = None
        for ext in (b"FIT_METADATA", b"STACK_DELTAS"):
            hit = [h for h in range(n) if strs[h, 0] == ext and kind[h, 1] == 1]
            if hit:
                order = int(inum[hit[0], 1])
                break
        else:
            userprint("WARNING: FITORDER not found in the delta attributes file")
            order = from_config()
    userprint(f"Setting order={order} for the polynomial used for the continuum fitting")
    return order


def _cosmo_tables(cosmo):
    """(z, r_comov, dist_m) tables behind ``cosmo.get_r_comov`` / ``cosmo.get_dist_m``: scipy
    ``interp1d`` objects in the reference (constants.py:211-229), ``table()`` on our own class."""
    if hasattr(cosmo, "table"):
        return tuple(np.ascontiguousarray(t, dtype=np.float64) for t in cosmo.table())
    f_r, f_m = cosmo.get_r_comov, cosmo.get_dist_m
    if not (hasattr(f_r, "x") and hasattr(f_r, "y") and hasattr(f_m, "x") and hasattr(f_m, "y")):
        raise TypeError("picca_b200.io: cannot find the distance tables of this cosmology object")
    if not np.array_equal(f_r.x, f_m.x):
        raise TypeError("picca_b200.io: r_comov and dist_m tables use different redshift grids")
    return (np.ascontiguousarray(f_r.x, dtype=np.float64),
            np.ascontiguousarray(f_r.y, dtype=np.float64),
            np.ascontiguousarray(f_m.y, dtype=np.float64))


class TableCosmo:
    """The distance tables of a cosmology object, detached from it (picklable): what the
    isolated loader process receives instead of the caller's ``Cosmo``."""

    def __init__(self, cosmo):
        self._tables = _cosmo_tables(cosmo)

    def table(self):
        return self._tables


def read_deltas(in_dir, nside, lambda_abs, alpha, z_ref, cosmo, max_num_spec=None,
                no_project=False, nproc=None, rebin_factor=None, z_min_qso=0, z_max_qso=10,
                delta_attributes=None, isolate=None):
    """Reads deltas and computes their redshifts, distances, evolved weights and projection
    (io.py:383-512).  Same arguments; returns ``(data, num_data, z_min, z_max)``.

    ``isolate`` (default: env ``PICCA_B200_IO_ISOLATE`` == "1"): run the device part in a
    short-lived child process (``python -m picca_b200._io_worker``) and take the SoA back through
    /dev/shm, so that THIS process never initialises CUDA.  The reference's scripts call
    ``read_deltas`` in the parent and only then fork their worker pool (picca_cf.py:387 then
    :455); a CUDA context cannot be used across a fork, so under the overlay the loader is
    isolated by default.

    Raises:
        AssertionError: if no healpix numbers are found (io.py:489-490)
        RuntimeError: projecting without a continuum order (data.py:628-633); CUDA errors
        ValueError: a redshift outside the cosmology table (scipy interp1d bounds error)
    """
    if isolate is None:
        isolate = os.environ.get("PICCA_B200_IO_ISOLATE", "0") == "1"
    kwds = dict(max_num_spec=max_num_spec, no_project=no_project, nproc=nproc,
                rebin_factor=rebin_factor, z_min_qso=z_min_qso, z_max_qso=z_max_qso,
                delta_attributes=delta_attributes)
    if isolate:
        soa = _read_soa_isolated(in_dir, nside, lambda_abs, alpha, z_ref,
                                 None if cosmo is None else TableCosmo(cosmo), **kwds)
    else:
        soa = read_deltas_soa(in_dir, nside, lambda_abs, alpha, z_ref, cosmo, **kwds)
    return soa_to_objects(soa)


def _read_soa_isolated(*args, **kwds):
    """``read_deltas_soa`` in a child process; arrays come back as .npy files in /dev/shm."""
    import pickle
    import shutil
    import subprocess
    import tempfile
    base = "/dev/shm" if os.path.isdir("/dev/shm") else None
    tmp = tempfile.mkdtemp(prefix="pb2_loader_", dir=base)
    try:
        with open(os.path.join(tmp, "args.pkl"), "wb") as f:
            pickle.dump((args, kwds), f)
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        env = dict(os.environ)
        env["PYTHONPATH"] = root + os.pathsep + env.get("PYTHONPATH", "")
        res = subprocess.run([sys.executable, "-m", "picca_b200._io_worker", tmp], env=env)
        meta_path = os.path.join(tmp, "meta.pkl")
        if not os.path.exists(meta_path):
            raise RuntimeError("picca_b200.io: the loader process died (exit code %d)"
                               % res.returncode)
        with open(meta_path, "rb") as f:
            meta = pickle.load(f)
        if "error" in meta:
            kind, msg = meta["error"]
            exc = {"AssertionError": AssertionError, "ValueError": ValueError,
                   "NotImplementedError": NotImplementedError, "OSError": OSError,
                   "TypeError": TypeError}.get(kind, RuntimeError)
            raise exc(msg)
        soa = dict(meta["scalars"])
        for name in meta["arrays"]:
            soa[name] = np.load(os.path.join(tmp, name + ".npy"))
        return soa
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def read_deltas_soa(in_dir, nside, lambda_abs, alpha, z_ref, cosmo, max_num_spec=None,
                    no_project=False, nproc=None, rebin_factor=None, z_min_qso=0, z_max_qso=10,
                    delta_attributes=None):
    """The loader proper: the forests of ``in_dir`` as ONE structure of host arrays (CSR
    ``offset`` + per-pixel ``log_lambda, delta, weights, z, r_comov, dist_m`` + per-forest
    metadata + ``healpix``), in file order.  ``read_deltas`` wraps it into the reference's data
    model; ``picca_b200.catalog.pack_soa`` packs it for the pair kernels without that detour."""
    in_dir = os.path.expandvars(in_dir)
    if len(in_dir) > 8 and in_dir[-8:] == '.fits.gz':
        files = sorted(glob.glob(in_dir))
    elif len(in_dir) > 5 and in_dir[-5:] == '.fits':
        files = sorted(glob.glob(in_dir))
    else:
        files = sorted(glob.glob(in_dir + '/*.fits') + glob.glob(in_dir + '/*.fits.gz'))
    order = find_order(in_dir, delta_attributes)

    eng = get_engine()  # raises without a device: no CPU fallback
    import time
    marks = [("start", time.perf_counter())]
    mark = lambda name: marks.append((name, time.perf_counter()))
    torch = eng.torch
    workers = nproc if nproc else (os.cpu_count() or 1)
    with ThreadPoolExecutor(max_workers=max(1, min(workers, 32))) as pool:
        parts = list(pool.map(lambda f: _open_delta_file(f, z_min_qso, z_max_qso,
                                                         rebin=rebin_factor is not None), files))
    for p in parts:  # ImageHDU files: the pixel counts come from the device
        if p.is_image:
            p.count_pixels(eng)

    mark("read + gunzip + header scan (host, %d threads)" % max(1, min(workers, 32)))
    # truncate like io.py:467-480: files are consumed in order until max_num_spec is exceeded
    if max_num_spec is not None:
        kept, total = [], 0
        for p in parts:
            kept.append(p)
            total += p.n
            if total > max_num_spec:
                break
        parts = kept
    n_los = sum(p.n for p in parts)
    if max_num_spec is not None:
        n_los = min(n_los, max_num_spec)
    if n_los == 0:
        raise AssertionError('ERROR: No data in {}'.format(in_dir))  # io.py:489-490

    cat = lambda name, dt: np.concatenate([np.asarray(getattr(p, name)) for p in parts]
                                          )[:n_los].astype(dt)
    ra, dec, z_qso = cat("ra", np.float64), cat("dec", np.float64), cat("z_qso", np.float64)
    los_id, plate = cat("los_id", np.int64), cat("plate", np.int64)
    mjd, fiberid = cat("mjd", np.int64), cat("fiberid", np.int64)
    n_pix = cat("n_pix", np.int64)
    wave_is_lambda = cat("wave_is_lambda", bool)
    if wave_is_lambda.any() and not wave_is_lambda.all():
        raise NotImplementedError("picca_b200.io: mixed LOGLAM / LAMBDA delta files")
    if not no_project and order is None:
        raise RuntimeError("Trying to project but order is not defined for the deltas. "
                           "Check previous warning to solve this issue")  # data.py:628-633
    if np.any(n_pix == 0):  # the reference's `z.min()` of an empty forest (io.py:497)
        raise ValueError("zero-size array to reduction operation minimum which has no identity")
    offset = np.zeros(n_los + 1, dtype=np.int64)
    np.cumsum(n_pix, out=offset[1:])
    total_pix = int(offset[-1])

    dev = eng.device
    f64 = lambda n: torch.empty(n, dtype=torch.float64, device=dev)
    d_ll, d_delta, d_w = f64(total_pix), f64(total_pix), f64(total_pix)
    d_offset = torch.from_numpy(offset).to(dev)

    # ---- H2D of the raw file bytes in batches + unpack at the CSR offsets
    def flush(batch, first):
        if not batch:
            return
        sizes = [p.buf.size for p in batch]
        base = np.concatenate([[0], np.cumsum(sizes)])
        staged = torch.empty(int(base[-1]), dtype=torch.uint8).pin_memory()
        host = staged.numpy()
        for p, b in zip(batch, base[:-1]):
            host[b:b + p.buf.size] = p.buf
        d_raw = staged.to(dev, non_blocking=True)
        n_b = min(sum(p.n for p in batch), n_los - first)
        row0 = np.concatenate([p.row0 + b for p, b in zip(batch, base[:-1])])[:n_b]
        row_bytes = np.concatenate([p.row_bytes for p in batch])[:n_b]
        col_off = np.concatenate([p.col_off for p in batch])[:n_b]
        d_row0 = torch.from_numpy(np.ascontiguousarray(row0, dtype=np.int64)).to(dev)
        d_rb = torch.from_numpy(np.ascontiguousarray(row_bytes, dtype=np.int32)).to(dev)
        d_co = torch.from_numpy(np.ascontiguousarray(col_off, dtype=np.int32)).to(dev)
        _lib.check(eng.lib.pb2_delta_unpack(
            ctypes.c_int64(n_b), ctypes.c_void_p(d_raw.data_ptr()),
            ctypes.c_void_p(d_row0.data_ptr()), ctypes.c_void_p(d_rb.data_ptr()),
            ctypes.c_void_p(d_co.data_ptr()),
            ctypes.c_void_p(d_offset.data_ptr() + 8 * first), ctypes.c_void_p(d_ll.data_ptr()),
            ctypes.c_void_p(d_delta.data_ptr()), ctypes.c_void_p(d_w.data_ptr()),
            eng.stream_ptr()), "pb2_delta_unpack")
        torch.cuda.current_stream().synchronize()  # the staging buffers die with this scope

    batch, batch_bytes, first, done = [], 0, 0, 0
    for p in parts:
        if done >= n_los:
            break
        if p.is_image:
            flush(batch, first)
            first, batch, batch_bytes = min(done, n_los), [], 0
            p.unpack(eng, min(p.n, n_los - first), d_offset.data_ptr() + 8 * first, d_ll, d_delta,
                     d_w)
            done += p.n
            first = min(done, n_los)
            continue
        batch.append(p)
        batch_bytes += p.buf.size
        done += p.n
        if batch_bytes >= _UPLOAD_BATCH:
            flush(batch, first)
            first, batch, batch_bytes = min(done, n_los), [], 0
    flush(batch, first)
    for p in parts:
        p.buf = None
        if p.is_image:
            p.d_raw = p.d_rows = None

    mark("pinned staging + H2D + unpack kernels")
    device_log10 = bool(wave_is_lambda.any())
    host_pow = os.environ.get("PICCA_B200_HOST_POW", "0") == "1"

    def host_log_lambda():
        """parity mode: log_lambda (log10 taken with NumPy when the array holds wavelengths)"""
        ll = d_ll.cpu().numpy()
        if device_log10:
            ll = np.log10(ll)
            d_ll.copy_(torch.from_numpy(ll))
        return ll

    # ---- Delta.rebin (data.py:657-686, io.py:362-378)
    if rebin_factor is not None:
        userprint(f"Rebinning deltas by a factor of {rebin_factor}\n")
        d_dwave = torch.from_numpy(np.concatenate(
            [np.full(p.n, p.dwave, dtype=np.float64) for p in parts])[:n_los].copy()).to(dev)
        d_wave = f64(total_pix)
        if host_pow or os.environ.get("PICCA_B200_REBIN_DEVICE_WAVE", "0") != "1":
            # data.py:666 with NumPy's power, always: the reference's bin count
            # ceil(((wave[-1] - wave[0]) / dwave + 1) / factor) sits exactly on an integer whenever
            # the forest length + 1 is a multiple of the factor, so one ulp in `wave` decides
            # whether a trailing bin exists; taking the reference's own power keeps its answer
            d_wave.copy_(torch.from_numpy(10**host_log_lambda()))
            device_log10 = False  # d_ll now holds log10 values
        else:
            _lib.check(eng.lib.pb2_delta_wave(
                ctypes.c_int64(total_pix), ctypes.c_int32(int(device_log10)),
                ctypes.c_void_p(d_ll.data_ptr()), ctypes.c_void_p(d_wave.data_ptr()),
                eng.stream_ptr()), "pb2_delta_wave")
        d_count = torch.zeros(n_los, dtype=torch.int32, device=dev)
        d_rstat = torch.zeros(1, dtype=torch.int32, device=dev)

        def rebin_pass(new_offset, outs):
            _lib.check(eng.lib.pb2_delta_rebin(
                ctypes.c_int64(n_los), ctypes.c_void_p(d_offset.data_ptr()),
                ctypes.c_void_p(d_wave.data_ptr()), ctypes.c_void_p(d_delta.data_ptr()),
                ctypes.c_void_p(d_w.data_ptr()), ctypes.c_void_p(d_dwave.data_ptr()),
                ctypes.c_int32(int(rebin_factor)), ctypes.c_void_p(d_count.data_ptr()),
                ctypes.c_void_p(d_rstat.data_ptr()),
                ctypes.c_void_p(new_offset.data_ptr()) if new_offset is not None else None,
                *[ctypes.c_void_p(t.data_ptr()) if t is not None else None for t in outs],
                eng.stream_ptr()), "pb2_delta_rebin")

        rebin_pass(None, (None, None, None))
        if int(d_rstat.item()) == 2:
            raise NotImplementedError("picca_b200.io: rebinning needs ascending wavelengths in "
                                      "every forest")
        n_pix = d_count.cpu().numpy().astype(np.int64)
        if np.any(n_pix == 0):  # the reference's `z.min()` of an emptied forest (io.py:497)
            raise ValueError("zero-size array to reduction operation minimum which has no "
                             "identity")
        offset = np.zeros(n_los + 1, dtype=np.int64)
        np.cumsum(n_pix, out=offset[1:])
        total_pix = int(offset[-1])
        d_new_offset = torch.from_numpy(offset).to(dev)
        new_ll, new_delta, new_w = f64(total_pix), f64(total_pix), f64(total_pix)
        rebin_pass(d_new_offset, (new_ll, new_delta, new_w))
        # the rebinned forests hold WAVELENGTHS until log10 is taken below (data.py:684)
        d_ll, d_delta, d_w, d_offset = new_ll, new_delta, new_w, d_new_offset
        device_log10 = True
        mark("rebin kernels")
    # ---- z, distances, weight evolution, projection (io.py:493-507)
    d_z, d_range = f64(total_pix), f64(2 * n_los)
    d_status = torch.zeros(1, dtype=torch.int32, device=dev)
    d_order = torch.full((n_los,), -1 if order is None else int(order), dtype=torch.int32,
                         device=dev)
    if cosmo is not None:
        tz, tr, tm = _cosmo_tables(cosmo)
        d_tz, d_tr, d_tm = (torch.from_numpy(t).to(dev) for t in (tz, tr, tm))
        d_rc, d_dm = f64(total_pix), f64(total_pix)
        tabs = (ctypes.c_int32(len(tz)), ctypes.c_void_p(d_tz.data_ptr()),
                ctypes.c_void_p(d_tr.data_ptr()), ctypes.c_void_p(d_tm.data_ptr()))
        dist = (ctypes.c_void_p(d_rc.data_ptr()), ctypes.c_void_p(d_dm.data_ptr()))
    else:  # io.py:500: distances only `if not cosmo is None`
        d_rc = d_dm = None
        tabs = (ctypes.c_int32(0), None, None, None)
        dist = (None, None)
    d_z_in = None
    if host_pow:
        ll_host = host_log_lambda()
        device_log10 = False
        d_z_in = torch.from_numpy(10**ll_host / lambda_abs - 1.).to(dev)  # io.py:496, NumPy power
    _lib.check(eng.lib.pb2_delta_prepare(
        ctypes.c_int64(n_los), ctypes.c_void_p(d_offset.data_ptr()),
        ctypes.c_void_p(d_order.data_ptr()), ctypes.c_double(lambda_abs), ctypes.c_double(alpha),
        ctypes.c_double(z_ref), *tabs, ctypes.c_int32(0 if no_project else 1),
        ctypes.c_int32(int(device_log10)),
        ctypes.c_void_p(d_z_in.data_ptr()) if d_z_in is not None else None,
        ctypes.c_void_p(d_ll.data_ptr()), ctypes.c_void_p(d_delta.data_ptr()),
        ctypes.c_void_p(d_w.data_ptr()), ctypes.c_void_p(d_z.data_ptr()), *dist,
        ctypes.c_void_p(d_range.data_ptr()), ctypes.c_void_p(d_status.data_ptr()),
        eng.stream_ptr()), "pb2_delta_prepare")
    if int(d_status.item()):
        raise ValueError("A value in x_new is outside the interpolation range of the cosmology "
                         "table (scipy interp1d bounds error in the reference)")

    eng.torch.cuda.synchronize()
    mark("prepare kernel")
    # ---- catalogue order (ascending HEALPix, file order inside a pixel: the iteration order of
    # fill_neighs, cf.py:91-122) before the copy back, so that catalog.pack can take the arrays
    # as they are; the gather runs in HBM
    healpixs = ang2pix_ring(nside, np.pi / 2. - dec, ra)  # io.py:486-488
    perm = np.argsort(healpixs, kind="stable")
    fields = [("log_lambda", d_ll), ("delta", d_delta), ("weights", d_w), ("z", d_z)]
    if d_rc is not None:
        fields += [("r_comov", d_rc), ("dist_m", d_dm)]
    z_range = d_range.cpu().numpy().reshape(n_los, 2)
    if not np.array_equal(perm, np.arange(n_los)):
        lengths = np.diff(offset)[perm]
        new_offset = np.zeros(n_los + 1, dtype=np.int64)
        np.cumsum(lengths, out=new_offset[1:])
        d_shift = torch.from_numpy(np.repeat(offset[:-1][perm] - new_offset[:-1], lengths)).to(dev)
        d_idx = d_shift + torch.arange(total_pix, dtype=torch.int64, device=dev)
        fields = [(name, t.index_select(0, d_idx)) for name, t in fields]
        offset = new_offset
        ra, dec, z_qso, los_id, plate, mjd, fiberid, healpixs = (
            v[perm] for v in (ra, dec, z_qso, los_id, plate, mjd, fiberid, healpixs))
    h = {name: t.cpu().numpy() for name, t in fields}
    z_min = float(z_range[:, 0].min())
    z_max = max(0., float(z_range[:, 1].max()))  # io.py:493: z_max starts at 0
    mark("D2H")
    if os.environ.get("PICCA_B200_IO_TIMING", "0") == "1":
        for (_, t0), (name, t1) in zip(marks[:-1], marks[1:]):
            userprint("picca_b200.io: %-52s %.3f s" % (name, t1 - t0))
    soa = dict(h)
    soa.update(offset=offset, ra=ra, dec=dec, z_qso=z_qso, los_id=los_id, plate=plate, mjd=mjd,
               fiberid=fiberid, healpix=np.asarray(healpixs, dtype=np.int64), n_los=int(n_los),
               z_min=z_min, z_max=z_max, order=order)
    return soa


SOA_ARRAYS = ("log_lambda", "delta", "weights", "z", "r_comov", "dist_m", "offset", "ra", "dec",
              "z_qso", "los_id", "plate", "mjd", "fiberid", "healpix")


def soa_to_objects(soa):
    """The reference's data model from the loader's SoA: ``dict[healpix] -> list[Delta]`` whose
    array attributes are views into the SoA (io.py:485-512 builds the same dict per forest).
    The SoA is in catalogue order and registered (``forest.register_soa``) so that
    ``catalog.pack`` takes the arrays as they are (after checking that the objects still are
    those views) instead of concatenating 100 000 per-forest arrays again."""
    offset, order = soa["offset"], soa["order"]
    has_dist = "r_comov" in soa
    data = {}
    los_id, ra, dec, z_qso = soa["los_id"], soa["ra"], soa["dec"], soa["z_qso"]
    plate, mjd, fiberid, healpixs = soa["plate"], soa["mjd"], soa["fiberid"], soa["healpix"]
    for f in range(soa["n_los"]):
        a, b = offset[f], offset[f + 1]
        d = Delta(int(los_id[f]), float(ra[f]), float(dec[f]), float(z_qso[f]), int(plate[f]),
                  int(mjd[f]), int(fiberid[f]), soa["log_lambda"][a:b], soa["weights"][a:b],
                  soa["delta"][a:b], order)
        d.z = soa["z"][a:b]
        if has_dist:
            d.r_comov, d.dist_m = soa["r_comov"][a:b], soa["dist_m"][a:b]
        data.setdefault(int(healpixs[f]), []).append(d)
    userprint("\n")
    register_soa(data, {k: soa[k] for k in ("offset",) + PIXEL_FIELDS if k in soa})
    return data, soa["n_los"], soa["z_min"], soa["z_max"]
