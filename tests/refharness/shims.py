"""Test-only shims that let the UNMODIFIED reference (``/root/reference/py/picca``) import and run
in this container, where healpy / fitsio / astropy / iminuit / camb / h5py are not installed
(SURVEY.md section 8c, Appendix D).  Nothing here is used by the product path.

* ``healpy.ang2pix`` (RING) follows the published HEALPix RING formulae; it reproduces the
  ``HEALPID`` column of the reference's golden files.
* ``healpy.query_disc(..., inclusive=True)`` returns an ascending SUPERSET of the true disc.
  That is equivalent for the reference because an exact ``ang < ang_max`` filter follows
  (reference ``py/picca/cf.py:123-125``, ``xcf.py:99-100``).
"""
import os
import sys
import types

import numpy as np

from . import minifits

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def _locate():
    """The reference tree: /root/reference in the build container, else the unmodified copy that
    ``scripts/stage_reference.py`` installed under the git-ignored ``baseline/_ref`` (it travels
    to the GPU box with the gpurun snapshot).  Returns (python path entry, test-data dir)."""
    if os.environ.get("PICCA_B200_FORCE_STAGED_REF", "0") != "1" and \
            os.path.isdir("/root/reference/py/picca"):
        return "/root/reference/py", "/root/reference/py/picca/tests/data"
    staged = os.path.join(_ROOT, "baseline", "_ref")
    if os.path.isfile(os.path.join(staged, "picca", "cf.py")):
        return staged, os.path.join(staged, "picca", "tests", "data")
    return None, None


REFERENCE_PY, REFERENCE_DATA = _locate()


def reference_available():
    return REFERENCE_PY is not None


def in_build_container():
    return REFERENCE_PY == "/root/reference/py"


# ----------------------------------------------------------------------------- HEALPix (RING)
def ang2pix(nside, theta, phi, nest=False, lonlat=False):
    assert not nest and not lonlat
    theta = np.asarray(theta, dtype=np.float64)
    phi = np.asarray(phi, dtype=np.float64)
    scalar = theta.ndim == 0
    theta = np.atleast_1d(theta)
    phi = np.atleast_1d(phi)
    z = np.cos(theta)
    za = np.abs(z)
    tt = np.mod(phi, 2.0 * np.pi) / (0.5 * np.pi)  # in [0,4)
    pix = np.empty(z.shape, dtype=np.int64)

    eq = za <= 2.0 / 3.0
    # equatorial belt
    temp1 = nside * (0.5 + tt[eq])
    temp2 = nside * z[eq] * 0.75
    jp = np.floor(temp1 - temp2).astype(np.int64)
    jm = np.floor(temp1 + temp2).astype(np.int64)
    ir = nside + 1 + jp - jm
    kshift = 1 - (ir & 1)
    ip = (jp + jm - nside + kshift + 1) // 2
    ip = np.mod(ip, 4 * nside)
    pix[eq] = 2 * nside * (nside - 1) + (ir - 1) * 4 * nside + ip

    # polar caps
    cap = ~eq
    tp = tt[cap] - np.floor(tt[cap])
    tmp = nside * np.sqrt(3.0 * (1.0 - za[cap]))
    jp = np.floor(tp * tmp).astype(np.int64)
    jm = np.floor((1.0 - tp) * tmp).astype(np.int64)
    ir = jp + jm + 1
    ip = np.floor(tt[cap] * ir).astype(np.int64)
    ip = np.mod(ip, 4 * ir)
    north = z[cap] > 0
    pix_cap = np.where(north, 2 * ir * (ir - 1) + ip, 12 * nside * nside - 2 * ir * (ir + 1) + ip)
    pix[cap] = pix_cap
    return int(pix[0]) if scalar else pix


def pix2vec(nside, pix):
    """Centres of RING pixels (unit vectors), standard HEALPix pix2ang_ring."""
    pix = np.asarray(pix, dtype=np.int64)
    npix = 12 * nside * nside
    ncap = 2 * nside * (nside - 1)
    z = np.empty(pix.shape)
    phi = np.empty(pix.shape)

    north = pix < ncap
    p = pix[north]
    iring = (1 + np.sqrt(1 + 2 * p).astype(np.int64)) >> 1
    # guard isqrt rounding
    iring = np.where(2 * iring * (iring - 1) > p, iring - 1, iring)
    iring = np.where(2 * iring * (iring + 1) <= p, iring + 1, iring)
    iphi = p + 1 - 2 * iring * (iring - 1)
    z[north] = 1.0 - iring**2 / (3.0 * nside * nside)
    phi[north] = (iphi - 0.5) * np.pi / (2.0 * iring)

    eq = (pix >= ncap) & (pix < npix - ncap)
    ip = pix[eq] - ncap
    iring = ip // (4 * nside) + nside
    iphi = ip % (4 * nside) + 1
    fodd = np.where(((iring + nside) & 1) == 1, 1.0, 0.5)
    z[eq] = (2 * nside - iring) * 2.0 / (3.0 * nside)
    phi[eq] = (iphi - fodd) * np.pi / (2.0 * nside)

    south = pix >= npix - ncap
    ip = npix - pix[south]
    iring = (1 + np.sqrt(2 * ip - 1).astype(np.int64)) >> 1
    iring = np.where(2 * iring * (iring - 1) >= ip, iring - 1, iring)
    iring = np.where(2 * iring * (iring + 1) < ip, iring + 1, iring)
    iphi = 4 * iring + 1 - (ip - 2 * iring * (iring - 1))
    z[south] = -1.0 + iring**2 / (3.0 * nside * nside)
    phi[south] = (iphi - 0.5) * np.pi / (2.0 * iring)

    st = np.sqrt(np.clip(1.0 - z * z, 0.0, None))
    return np.stack([st * np.cos(phi), st * np.sin(phi), z], axis=-1)


_CENTRES = {}


def query_disc(nside, vec, radius, inclusive=False, fact=4, nest=False):
    """Ascending superset of the pixels overlapping the disc (margin = 2 pixel sizes)."""
    if nside not in _CENTRES:
        _CENTRES[nside] = pix2vec(nside, np.arange(12 * nside * nside))
    centres = _CENTRES[nside]
    vec = np.asarray(vec, dtype=np.float64)
    vec = vec / np.sqrt(np.sum(vec * vec))
    margin = 2.0 * np.sqrt(4.0 * np.pi / (12.0 * nside * nside))
    lim = np.cos(min(np.pi, radius + margin))
    return np.nonzero(centres @ vec >= lim)[0]


# ----------------------------------------------------------------------------- astropy.table.Table
class Table:
    """The few Table features ``io.read_drq`` / ``io.read_objects`` use."""

    def __init__(self, data=None):
        self._cols = {}
        if isinstance(data, np.ndarray) and data.dtype.names:
            for name in data.dtype.names:
                self._cols[name] = np.array(data[name])
        elif isinstance(data, dict):
            for name, val in data.items():
                self._cols[name] = np.array(val)

    @property
    def colnames(self):
        return list(self._cols)

    def rename_column(self, old, new):
        self._cols = {(new if k == old else k): v for k, v in self._cols.items()}

    def keep_columns(self, names):
        self._cols = {k: v for k, v in self._cols.items() if k in names}

    def __len__(self):
        return len(next(iter(self._cols.values()))) if self._cols else 0

    def __getitem__(self, item):
        if isinstance(item, str):
            return self._cols[item]
        if isinstance(item, (int, np.integer)):
            return {k: v[item] for k, v in self._cols.items()}
        return Table({k: v[item] for k, v in self._cols.items()})

    def __setitem__(self, key, value):
        self._cols[key] = np.asarray(value)

    def __iter__(self):
        for k in range(len(self)):
            yield self[k]


# ----------------------------------------------------------------------------- installation
def install():
    """Register the stub modules and put the reference on ``sys.path``.  Returns True when the
    reference tree is present (it is absent on the GPU box)."""
    if not reference_available():
        return False
    if "healpy" not in sys.modules:
        healpy = types.ModuleType("healpy")
        healpy.ang2pix = ang2pix
        healpy.query_disc = query_disc
        sys.modules["healpy"] = healpy
    if "fitsio" not in sys.modules:
        fitsio = types.ModuleType("fitsio")
        fitsio.FITS = minifits.FITS
        fitsio.read = minifits.read
        sys.modules["fitsio"] = fitsio
    if "astropy" not in sys.modules:
        astropy = types.ModuleType("astropy")
        table = types.ModuleType("astropy.table")
        table.Table = Table
        astropy.table = table
        sys.modules["astropy"] = astropy
        sys.modules["astropy.table"] = table
    for name in ("iminuit", "camb", "h5py"):
        if name not in sys.modules:
            mod = types.ModuleType(name)
            if name == "iminuit":
                mod.Minuit = object
            sys.modules[name] = mod
    if REFERENCE_PY not in sys.path:
        sys.path.insert(0, REFERENCE_PY)
    return True
