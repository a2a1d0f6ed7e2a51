"""Host-side (NumPy) restatement of the reference's per-forest Python logic.
TEST INFRASTRUCTURE ONLY.

  angle_between_many / angle_between_one <- data.QSO.get_angle_between (py/picca/data.py:106-162)
  same_half_plate                        <- cf.py:171-185
  Catalogue                              <- the iteration order of fill_neighs (cf.py:91-122):
                                            ascending HEALPix id, then list order.
"""
import numpy as np

SMALL_ANGLE_CUT_OFF = 2. / 3600. * np.pi / 180.  # constants.py:16
SPEED_LIGHT = 299792458.0 / 1000.  # constants.py:18 [km/s]


class Catalogue:
    """Flat view of ``dict[healpix] -> list[obj]`` in ascending-healpix, list order."""

    def __init__(self, data):
        self.healpixs = sorted(data)
        self.objs = [obj for hp in self.healpixs for obj in data[hp]]
        self.first = {}
        k = 0
        for hp in self.healpixs:
            self.first[hp] = k
            k += len(data[hp])
        self.x = np.array([o.x_cart for o in self.objs], dtype=np.float64)
        self.y = np.array([o.y_cart for o in self.objs], dtype=np.float64)
        self.z = np.array([o.z_cart for o in self.objs], dtype=np.float64)
        self.ra = np.array([o.ra for o in self.objs], dtype=np.float64)
        self.dec = np.array([o.dec for o in self.objs], dtype=np.float64)
        self.thingid = np.array([o.thingid for o in self.objs])


_CACHE = {}


def catalogue(data):
    key = id(data)
    hit = _CACHE.get(key)
    if hit is None or hit[0] is not data or len(hit[1].objs) != sum(len(v) for v in data.values()):
        _CACHE[key] = (data, Catalogue(data))
    return _CACHE[key][1]


def angle_between_many(obj, cat, sel=None):
    """get_angle_between, list-like branch (data.py:118-141), against catalogue rows ``sel``."""
    sl = slice(None) if sel is None else sel
    x_cart, y_cart, z_cart, ra, dec = cat.x[sl], cat.y[sl], cat.z[sl], cat.ra[sl], cat.dec[sl]
    cos = x_cart * obj.x_cart + y_cart * obj.y_cart + z_cart * obj.z_cart
    cos = np.where(cos >= 1., 1., cos)
    cos = np.where(cos <= -1., -1., cos)
    angl = np.arccos(cos)
    w = ((np.absolute(ra - obj.ra) < SMALL_ANGLE_CUT_OFF) &
         (np.absolute(dec - obj.dec) < SMALL_ANGLE_CUT_OFF))
    if w.sum() != 0:
        angl[w] = np.sqrt((dec[w] - obj.dec)**2 + (obj.cos_dec * (ra[w] - obj.ra))**2)
    return angl


def angle_between_one(obj, other):
    """get_angle_between, scalar branch (data.py:143-161)."""
    cos = other.x_cart * obj.x_cart + other.y_cart * obj.y_cart + other.z_cart * obj.z_cart
    if cos >= 1.:
        cos = 1.
    elif cos <= -1.:
        cos = -1.
    angl = np.arccos(cos)
    if ((np.absolute(other.ra - obj.ra) < SMALL_ANGLE_CUT_OFF) &
            (np.absolute(other.dec - obj.dec) < SMALL_ANGLE_CUT_OFF)):
        angl = np.sqrt((other.dec - obj.dec)**2 + (obj.cos_dec * (other.ra - obj.ra))**2)
    return angl


def same_half_plate(delta1, delta2):
    """cf.py:171-183 (incl. the RuntimeError for string fiberids)."""
    if isinstance(delta1.fiberid, str) or isinstance(delta2.fiberid, str):
        raise RuntimeError("Trying to figure out if two spectra "
                           "come from the same half plate but "
                           "combined reobservations were given")
    return bool((delta1.plate == delta2.plate) and
                ((delta1.fiberid <= 500 and delta2.fiberid <= 500) or
                 (delta1.fiberid > 500 and delta2.fiberid > 500)))


class _NoLock:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


class _Counter:
    value = 0


def progress(mod):
    """The progress counter side effect of cf.py:163-167 (kept, minus the print)."""
    lock = mod.lock if getattr(mod, "lock", None) is not None else _NoLock()
    counter = mod.counter if getattr(mod, "counter", None) is not None else _Counter
    with lock:
        counter.value += 1
