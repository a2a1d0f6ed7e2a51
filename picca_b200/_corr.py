"""Host logic shared by picca_b200.cf and picca_b200.xcf: neighbour batches, lazy neighbour
views, progress accounting.  No arithmetic of the hot path happens here."""
import os

import numpy as np

from . import catalog as _catalog
from .engine import get_engine

HOST_ANGLES = os.environ.get("PICCA_B200_HOST_ANGLES", "0") == "1"

# reference py/picca/constants.py:16
SMALL_ANGLE_CUT_OFF = 2. / 3600. * np.pi / 180.


class LazyNeighbours:
    """What ``delta.neighbours`` holds between fill_neighs and compute_*: a sequence view of the
    device neighbour list of one line of sight (reference stores a list/array of objects,
    cf.py:125-135, xcf.py:123).  Materialised only if somebody iterates it."""

    __slots__ = ("_pairs", "_k", "_objs2")

    def __init__(self, pairs, k, objs2):
        self._pairs, self._k, self._objs2 = pairs, k, objs2

    def _index(self):
        off = self._pairs.host_offset()
        return self._pairs.host_f2()[off[self._k]:off[self._k + 1]]

    def __len__(self):
        off = self._pairs.host_offset()
        return int(off[self._k + 1] - off[self._k])

    @property
    def size(self):
        return len(self)

    def __iter__(self):
        return (self._objs2[q] for q in self._index())

    def __getitem__(self, item):
        idx = self._index()[item]
        if np.ndim(idx) == 0:
            return self._objs2[int(idx)]
        return [self._objs2[q] for q in idx]


def set_neighbours(obj, value):
    """``obj.neighbours = value`` without going through a Python-level ``__setattr__`` (the
    forests of a registered catalogue watch their attributes, forest.Delta.__setattr__): this runs
    once per forest and per call."""
    try:
        obj.__dict__["neighbours"] = value
    except (AttributeError, TypeError):
        obj.neighbours = value


def set_lazy_neighbours(objs1, index, pairs, objs2):
    """``objs1[f1].neighbours = LazyNeighbours(pairs, k, objs2)`` for f1 = index[k]: once per
    forest and fill_neighs call (100 000 times on config 2), so without a function call per
    forest where the objects have an instance dict."""
    new = LazyNeighbours
    try:
        for k, f1 in enumerate(index):
            objs1[f1].__dict__["neighbours"] = new(pairs, k, objs2)
    except (AttributeError, TypeError):   # slotted / foreign classes: plain attribute assignment
        for k, f1 in enumerate(index):
            objs1[f1].neighbours = new(pairs, k, objs2)


def clear_neighbours(objs1, index):
    """``objs1[f1].neighbours = None`` for every f1 of ``index`` (what the reference does once a
    forest has been used, cf.py:240, xcf.py:213)."""
    try:
        for f1 in index:
            objs1[f1].__dict__["neighbours"] = None
    except (AttributeError, TypeError):
        for f1 in index:
            objs1[f1].neighbours = None


class PendingNeighbours:
    """``delta.neighbours`` after a DEFERRED fill_neighs (see ``defer_fill``): the neighbour
    search runs when somebody looks at the list or a compute_* function needs it."""

    def __init__(self, owner, fill_now, healpixs):
        self._owner, self._fill_now, self._healpixs = owner, fill_now, healpixs

    def _real(self):
        if self._owner.neighbours is self:
            self._fill_now(self._healpixs)     # replaces .neighbours of every forest of the list
        return self._owner.neighbours

    def __len__(self):
        return len(self._real())

    @property
    def size(self):
        return len(self)

    def __iter__(self):
        return iter(self._real())

    def __getitem__(self, item):
        return self._real()[item]


def defer_fill():
    """fill_neighs is deferred when it is called in the MAIN process before this process has
    touched CUDA: picca_xwick.py fills the neighbours in the parent and only then forks its pool
    (picca_xwick.py:446, :452), and a CUDA context does not survive a fork.  The forked workers
    (and an in-process caller such as picca_dmat.py --nproc 1) run the search on first use."""
    import multiprocessing
    from . import engine
    if engine._ENGINE is not None and engine._ENGINE._pid == os.getpid():
        return False
    if os.environ.get("PICCA_B200_EAGER_FILL", "0") == "1":
        return False
    return multiprocessing.current_process().name == "MainProcess"


class NeighbourStore:
    """Neighbour lists produced by fill_neighs, kept on the device until compute_* uses them."""

    def __init__(self):
        self.by_healpix = {}
        self.pending = set()      # healpixs whose fill_neighs was deferred

    def defer(self, healpixs, data, fill_now):
        healpixs = list(healpixs)
        for hp in healpixs:
            self.pending.add(hp)
            self.by_healpix.pop(hp, None)
            for obj in data[hp]:
                set_neighbours(obj, PendingNeighbours(obj, fill_now, healpixs))

    def put(self, healpixs, pairs, ranges, cats=()):
        """``cats``: the packed catalogues the list was built from (kept alive with it)."""
        for hp in healpixs:
            self.by_healpix[hp] = (pairs, ranges[hp], tuple(cats))
            self.pending.discard(hp)

    def take(self, healpixs, cats=()):
        """The stored PairList when ``healpixs`` is exactly one stored batch built from the
        catalogues ``cats``; None when it has to be rebuilt (stored in other batches, or built
        from a catalogue that has been re-packed since)."""
        missing = [hp for hp in healpixs if hp not in self.by_healpix]
        if missing:
            if all(hp in self.pending for hp in missing):
                return None   # deferred fill_neighs: the caller runs it now
            raise RuntimeError("picca_b200: compute called before fill_neighs for healpix %r"
                               % (missing[:5],))
        for hp in healpixs:
            have = self.by_healpix[hp][2]
            if len(have) != len(cats) or any(a is not b for a, b in zip(have, cats)):
                return None
        first = self.by_healpix[healpixs[0]][0]
        same = all(self.by_healpix[hp][0] is first for hp in healpixs)
        covered = sum(e[1][1] - e[1][0] for e in (self.by_healpix[hp] for hp in healpixs))
        in_order = same and all(
            self.by_healpix[a][1][1] == self.by_healpix[b][1][0]
            for a, b in zip(healpixs[:-1], healpixs[1:]))
        if same and in_order and covered == first.n_f1:
            return first
        return None

    def drop(self, healpixs):
        for hp in healpixs:
            self.by_healpix.pop(hp, None)
            self.pending.discard(hp)


def forest_index_of(host_cat, healpixs):
    """Catalogue indices of the lines of sight of ``healpixs`` (in call order) + per-healpix
    [k0, k1) ranges inside that list."""
    parts, ranges, k = [], {}, 0
    for hp in healpixs:
        a, b = host_cat.first_of(hp)
        parts.append(np.arange(a, b, dtype=np.int32))
        ranges[hp] = (k, k + (b - a))
        k += b - a
    index = np.concatenate(parts) if parts else np.zeros(0, dtype=np.int32)
    return index, ranges


def host_angles(cat1, cat2, f1_of_pair, f2_of_pair):
    """QSO.get_angle_between for every listed pair, evaluated on the host with NumPy exactly as
    the reference does (data.py:126-141); used by the parity mode only."""
    A, B = cat1.arrays, cat2.arrays
    cos = (B["x_cart"][f2_of_pair] * A["x_cart"][f1_of_pair] +
           B["y_cart"][f2_of_pair] * A["y_cart"][f1_of_pair] +
           B["z_cart"][f2_of_pair] * A["z_cart"][f1_of_pair])
    cos = np.where(cos >= 1., 1., cos)
    cos = np.where(cos <= -1., -1., cos)
    ang = np.arccos(cos)
    dra = B["ra"][f2_of_pair] - A["ra"][f1_of_pair]
    ddec = B["dec"][f2_of_pair] - A["dec"][f1_of_pair]
    w = (np.absolute(dra) < SMALL_ANGLE_CUT_OFF) & (np.absolute(ddec) < SMALL_ANGLE_CUT_OFF)
    if w.sum() != 0:
        ang[w] = np.sqrt(ddec[w]**2 + (A["cos_dec"][f1_of_pair][w] * dra[w])**2)
    return ang


def apply_host_angles(pairs, cat1, cat2):
    f1 = pairs.f1_index.cpu().numpy()[pairs.nb_f1.cpu().numpy()]
    pairs.set_host_angles(host_angles(cat1, cat2, f1, pairs.host_f2()))


class _NoLock:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


def bump_progress(mod, n_forests, userprint):
    """The shared progress counter of cf.py:163-167 / xcf.py:151-155, advanced in bulk."""
    counter = getattr(mod, "counter", None)
    if counter is None or n_forests == 0:
        return
    lock = getattr(mod, "lock", None) or _NoLock()
    num_data = getattr(mod, "num_data", None) or 1
    with lock:
        before = counter.value
        counter.value += n_forests
        if before // 1000 != counter.value // 1000 or before == 0:
            userprint("computing xi: {}%".format(round(before * 100.0 / num_data, 2)))


def engine_and_catalog(data, is_object=False, ang_correlation=False):
    eng = get_engine()
    host = _catalog.cached_pack(data, is_object=is_object, ang_correlation=ang_correlation)
    if id(host) not in eng._cats:  # a (re-)packed catalogue: release device copies of stale ones
        live = {id(hit[1]) for hit in _catalog._HOST_CACHE.values()}
        for key in [k for k in eng._cats if k not in live]:
            del eng._cats[key]
    return eng, host, eng.device_catalog(host)
