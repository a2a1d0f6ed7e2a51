"""B200 implementation of ``picca.co`` -- the correlation of two object catalogues -- behind the
reference's module API (reference py/picca/co.py): same module globals (:17-32, plus the
``z_cut_min`` / ``z_cut_max`` / ``num_data`` the script assigns, picca_co.py:220-257), same
functions and return tuple:

    fill_neighs(healpixs)                                   co.py:35-74
    compute_xi(healpixs) -> (weights, r_par, r_trans, z, num_pairs)   co.py:77-132

so that ``picca_co.py`` runs unchanged once this module is importable as ``picca.co``
(``picca_b200.overlay``).  The pair arithmetic runs in ``pb2_co_pairs`` (csrc/pb2_co.cu) on the
neighbour list of the device neighbour search; no Numba, no CPU fallback.
"""
import ctypes
import sys

import numpy as np

from . import _corr, _lib
from .engine import MODE_CROSS
from .params import params_from_module


def userprint(*args, **kwds):
    """reference py/picca/utils.py:20-28"""
    print(*args, **kwds)
    sys.stdout.flush()


# ---- module globals: names and defaults of reference co.py:17-32
num_bins_r_par = None
num_bins_r_trans = None
r_par_min = None
r_par_max = None
r_trans_max = None
ang_max = None
nside = None

objs = None
objs2 = None

type_corr = None
x_correlation = False

counter = None
lock = None

# assigned by picca_co.py (:220-221, :257) and read by fill_neighs / compute_xi
z_cut_min = None
z_cut_max = None
num_data = None

_THIS = sys.modules[__name__]
_STORE = _corr.NeighbourStore()


def _catalogs():
    eng, host1, dev1 = _corr.engine_and_catalog(objs, is_object=True)
    if objs2 is not None:
        _, host2, dev2 = _corr.engine_and_catalog(objs2, is_object=True)
    else:
        host2, dev2 = host1, dev1
    return eng, host1, dev1, host2, dev2


def fill_neighs(healpixs):
    """Neighbouring objects of every object of ``healpixs`` (co.py:35-74): other thingid, angle
    below ``ang_max`` (no ordering: an auto-correlation visits every pair from both ends, as the
    reference does); the mean-redshift cut of :70-74 is applied when the pairs are counted."""
    healpixs = list(healpixs)
    if _corr.defer_fill():   # main process, no CUDA yet: see _corr.defer_fill
        _STORE.defer(healpixs, objs, _fill_neighs_now)
        return
    _fill_neighs_now(healpixs)


def _fill_neighs_now(healpixs):
    healpixs = list(healpixs)
    eng, host1, dev1, host2, dev2 = _catalogs()
    params = params_from_module(_THIS)
    index, ranges = _corr.forest_index_of(host1, healpixs)
    pairs = eng.neighbours(dev1, dev2, params, MODE_CROSS, index)
    if _corr.HOST_ANGLES:
        _corr.apply_host_angles(pairs, host1, host2)
    _STORE.put(healpixs, pairs, ranges, (host1, host2))
    _corr.set_lazy_neighbours(host1.objs, index, pairs, host2.objs)


def compute_xi(healpixs):
    """Pair counts of the objects of ``healpixs`` with their neighbours (co.py:77-132).
    Returns (weights, r_par, r_trans, z, num_pairs); r_par, r_trans, z normalised by the weights
    per call (:128-131)."""
    healpixs = list(healpixs)
    eng, host1, dev1, host2, dev2 = _catalogs()
    params = params_from_module(_THIS)
    pairs = _STORE.take(healpixs, (host1, host2))
    if pairs is None:
        _fill_neighs_now(healpixs)
        pairs = _STORE.take(healpixs, (host1, host2))
    torch = eng.torch
    nb = params.num_bins_r_par * params.num_bins_r_trans
    out = torch.zeros((1, 5, nb), dtype=torch.float64, device=eng.device)
    out_row = torch.zeros(max(pairs.n_f1, 1), dtype=torch.int32, device=eng.device)
    take_abs = (not x_correlation) or type_corr in ['DR', 'RD']  # co.py:172
    has_cut = z_cut_min is not None and z_cut_max is not None
    ps = pairs.struct()
    _lib.check(eng.lib.pb2_co_pairs(
        ctypes.byref(dev1.struct), ctypes.byref(dev2.struct), ctypes.byref(params),
        ctypes.byref(ps), ctypes.c_int32(int(take_abs)), ctypes.c_int32(int(has_cut)),
        ctypes.c_double(z_cut_min if has_cut else 0.), ctypes.c_double(z_cut_max if has_cut else 0.),
        ctypes.c_void_p(out_row.data_ptr()), ctypes.c_int64(1), ctypes.c_void_p(out.data_ptr()),
        eng.stream_ptr()), "pb2_co_pairs")
    host = out.cpu().numpy()[0]
    _corr.bump_progress(_THIS, pairs.n_f1, userprint)
    # co.py:126: objects with at least one neighbour (after the redshift cut) drop their list;
    # the others are skipped at :107-108 and keep theirs
    offset = pairs.host_offset()
    f1_index = pairs.f1_index.cpu().numpy()
    lengths = np.diff(offset)
    has = lengths > 0
    if has_cut and offset[-1] > 0:
        z1 = np.repeat(host1.arrays["z_qso"][f1_index], lengths)
        zm = (host2.arrays["z_qso"][pairs.host_f2()] + z1) / 2.
        ok = ((zm >= z_cut_min) & (zm < z_cut_max)).astype(np.int64)
        has = np.where(has, np.add.reduceat(np.append(ok, 0), offset[:-1]) > 0, False)
    for k, f1 in enumerate(f1_index):
        if has[k]:
            _corr.set_neighbours(host1.objs[f1], None)
    _STORE.drop(healpixs)
    weights = np.ascontiguousarray(host[0])
    r_par, r_trans, z = (np.ascontiguousarray(host[k]) for k in (1, 2, 3))
    num_pairs = np.ascontiguousarray(host[4]).view(np.int64)
    w = weights > 0.
    r_par[w] /= weights[w]
    r_trans[w] /= weights[w]
    z[w] /= weights[w]
    return weights, r_par, r_trans, z, num_pairs
