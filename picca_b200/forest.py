"""Line-of-sight records with the attribute names the reference's data model uses.

The product consumes ``dict[healpix] -> list[Delta]`` / ``list[QSO]`` exactly as
``picca.io.read_deltas`` / ``read_objects`` produce them (reference ``py/picca/data.py:14-162`` for
``QSO``, ``:238-373`` for ``Delta``).  When the reference's own classes are available they are used
as-is (duck typing); these light-weight stand-ins carry the same attributes so that the synthetic
generator, the tests and the benchmark can build inputs where the reference is not installed.
"""
import numpy as np

# reference py/picca/constants.py:16 -- 2 arcsec
SMALL_ANGLE_CUT_OFF = 2. / 3600. * np.pi / 180.


class QSO:
    """Attributes as reference ``data.QSO`` (``py/picca/data.py:60-104``)."""

    def __init__(self, los_id, ra, dec, z_qso, plate, mjd, fiberid):
        self.ra = ra
        self.dec = dec
        self.plate = plate
        self.mjd = mjd
        self.fiberid = fiberid
        self.x_cart = np.cos(ra) * np.cos(dec)
        self.y_cart = np.sin(ra) * np.cos(dec)
        self.z_cart = np.sin(dec)
        self.cos_dec = np.cos(dec)
        self.z_qso = z_qso
        self.los_id = los_id
        self.thingid = los_id
        self.weights = None
        self.r_comov = None
        self.dist_m = None
        self.log_lambda = None
        self.neighbours = None


class SoAToken:
    """Shared by the forests a producer built as views into one SoA (see ``register_soa``):
    re-binding a per-pixel or positional attribute of any of them marks the whole catalogue
    dirty, which sends ``catalog.pack`` back to its generic (object-walking) path."""
    __slots__ = ("dirty",)

    def __init__(self):
        self.dirty = False


_WATCHED = frozenset(("log_lambda", "delta", "weights", "z", "r_comov", "dist_m", "ra", "dec",
                      "x_cart", "y_cart", "z_cart", "cos_dec", "z_qso", "thingid", "plate",
                      "fiberid", "order"))


class Delta(QSO):
    """Attributes as reference ``data.Delta`` (``py/picca/data.py:296-373``), hot-path subset."""
    _soa_token = None

    def __setattr__(self, name, value):
        object.__setattr__(self, name, value)
        if name in _WATCHED:
            token = self._soa_token
            if token is not None:
                token.dirty = True

    def __init__(self, los_id, ra, dec, z_qso, plate, mjd, fiberid, log_lambda, weights, delta,
                 order):
        QSO.__init__(self, los_id, ra, dec, z_qso, plate, mjd, fiberid)
        self.log_lambda = log_lambda
        self.weights = weights
        self.delta = delta
        self.order = order
        self.z = None
        self.r_comov = None
        self.dist_m = None
        self.neighbours = None


# ---------------------------------------------------------------------------------------------
# SoA registry: a producer that builds the forests of a ``data`` dict as VIEWS into one
# contiguous array per field, forests back to back in catalogue order (ascending HEALPix, list
# order), registers those arrays here; ``catalog.pack`` then packs straight from them instead of
# concatenating 100 000 per-forest arrays again -- after checking, pointer by pointer, that the
# objects still are those views (any re-bound attribute sends it back to the generic path).
PIXEL_FIELDS = ("log_lambda", "delta", "weights", "z", "r_comov", "dist_m")
_SOA_OF = {}


LOS_FIELDS = ("x_cart", "y_cart", "z_cart", "ra", "dec", "cos_dec", "z_qso", "thingid", "plate",
              "fiberid", "order")


def register_soa(data, soa):
    """``soa``: dict with ``offset`` (int64[n_los + 1]) and the PIXEL_FIELDS arrays (float64,
    C-contiguous, catalogue order).  The per-line-of-sight attributes are gathered from the
    objects here, once; the forests get a shared ``SoAToken`` so that a later re-binding of any
    watched attribute is noticed without walking 100 000 objects again (in-place edits of the
    arrays need no notice: the views share the SoA's memory)."""
    import numpy as np
    if len(_SOA_OF) > 8:
        _SOA_OF.clear()
    objs = [obj for hp in sorted(data) for obj in data[hp]]
    entry = dict(soa)
    entry["objs"] = objs
    watched = all(type(o) is Delta for o in objs)
    if watched and objs:
        los = {}
        for name in LOS_FIELDS:
            vals = [getattr(o, name) for o in objs]
            if name == "order":
                vals = [-1 if v is None else int(v) for v in vals]
            # numeric columns are kept as arrays (catalog.pack slices them on every re-pack;
            # converting 100 000-element lists again costs 5 ms per column and call); ids that
            # are not numbers (combined re-observations) stay lists
            try:
                arr = np.array(vals)
                if arr.dtype.kind in "iuf" and arr.shape == (len(vals),):
                    vals = arr
            except (TypeError, ValueError):
                pass
            los[name] = vals
        entry["los"] = los
        token = SoAToken()
        for o in objs:
            object.__setattr__(o, "_soa_token", token)
        entry["token"] = token
        entry["lists"] = tuple((hp, id(v), len(v)) for hp, v in sorted(data.items()))
    _SOA_OF[id(data)] = (data, entry)


def soa_of(data):
    hit = _SOA_OF.get(id(data))
    return hit[1] if hit is not None and hit[0] is data else None


def registered_clean(data, soa):
    """True when ``data`` still is exactly what its producer registered: same per-pixel list
    objects of the same lengths and no watched attribute re-bound since (``SoAToken``)."""
    token = soa.get("token")
    if token is None or token.dirty:
        return False
    return soa.get("lists") == tuple((hp, id(v), len(v)) for hp, v in sorted(data.items()))


def views_intact(objs, soa, fields):
    """True when attribute ``name`` of object k is ``soa[name][offset[k]:offset[k+1]]`` for every
    k and every name in ``fields``: a view of that very array (``.base``), of that length, and --
    checked address by address on the ``weights`` field, whose forests are laid out like all the
    others -- at that position."""
    import numpy as np
    offset = soa["offset"]
    n = len(objs)
    if n != len(offset) - 1:
        return False
    sizes = np.diff(offset)
    try:
        for name in fields:
            base = soa[name]
            if base.dtype != np.float64 or not base.flags.c_contiguous:
                return False
            root = base if base.base is None else base.base
            for o in objs:
                v = getattr(o, name)
                if v.base is not base and v.base is not root:
                    return False
            lens = np.fromiter((getattr(o, name).size for o in objs), dtype=np.int64, count=n)
            if not np.array_equal(lens, sizes):
                return False
        base = soa["weights"]
        ptr0 = base.__array_interface__["data"][0]
        got = np.fromiter((o.weights.__array_interface__["data"][0] for o in objs),
                          dtype=np.int64, count=n)
    except (AttributeError, TypeError):
        return False
    return bool(np.array_equal(got[sizes > 0], (ptr0 + 8 * offset[:-1])[sizes > 0]))
