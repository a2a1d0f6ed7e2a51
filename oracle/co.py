"""Oracle double of the reference module ``picca.co`` (object x object correlation): same
globals, same functions, same return tuple, NumPy on the CPU.
TEST INFRASTRUCTURE ONLY -- the referee for picca_b200.co, never the product.

Restates: fill_neighs (co.py:35-74), compute_xi (co.py:77-132), compute_xi_forest_pairs
(co.py:135-202).  Pinned bit for bit against the live reference on its bundled quasar catalogue
(tests/test_oracle_vs_reference.py) and against tests/golden/golden_co.npz.
"""
import sys

import numpy as np

from . import _host

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
z_cut_min = None
z_cut_max = None
num_data = None

_THIS = sys.modules[__name__]


def fill_neighs(healpixs):
    """co.py:35-74.  Candidates are all objects of the second catalogue (any superset of the
    disc query is equivalent: the exact angle filter follows, :67-69)."""
    cat = _host.catalogue(objs2 if objs2 is not None else objs)
    for healpix in healpixs:
        for obj1 in objs[healpix]:
            ang = _host.angle_between_many(obj1, cat)
            w = (cat.thingid != obj1.thingid) & (ang < ang_max)
            neighbours = [cat.objs[k] for k in np.nonzero(w)[0]]
            obj1.neighbours = np.array([
                obj2 for obj2 in neighbours
                if ((obj2.z_qso + obj1.z_qso) / 2. >= z_cut_min and
                    (obj2.z_qso + obj1.z_qso) / 2. < z_cut_max)])        # :70-74


class _Rows:
    """the attribute arrays QSO.get_angle_between gathers from a list of objects (data.py:119-123)"""

    def __init__(self, objects):
        self.x = np.array([o.x_cart for o in objects])
        self.y = np.array([o.y_cart for o in objects])
        self.z = np.array([o.z_cart for o in objects])
        self.ra = np.array([o.ra for o in objects])
        self.dec = np.array([o.dec for o in objects])


def compute_xi_forest_pairs(z1, r_comov1, dist_m1, weights1, z2, r_comov2, dist_m2, weights2, ang):
    """co.py:170-202"""
    r_par = (r_comov1 - r_comov2) * np.cos(ang / 2.)
    if not x_correlation or type_corr in ['DR', 'RD']:
        r_par = np.absolute(r_par)
    r_trans = (dist_m1 + dist_m2) * np.sin(ang / 2.)
    z = (z1 + z2) / 2.
    weights12 = weights1 * weights2
    w = ((r_par >= r_par_min) & (r_par < r_par_max) & (r_trans < r_trans_max) & (weights12 > 0.))
    r_par, r_trans, z, weights12 = r_par[w], r_trans[w], z[w], weights12[w]
    bins = (((r_par - r_par_min) / (r_par_max - r_par_min) * num_bins_r_par).astype(np.int64) *
            num_bins_r_trans + (r_trans / r_trans_max * num_bins_r_trans).astype(np.int64))
    nb = int(bins.max()) + 1 if bins.size else 0
    out = [np.zeros(nb), np.zeros(nb), np.zeros(nb), np.zeros(nb), np.zeros(nb, dtype=np.int64)]
    for target, values in zip(out, (weights12, r_par * weights12, r_trans * weights12,
                                    z * weights12, np.ones(bins.size, dtype=np.int64))):
        np.add.at(target, bins, values)   # numba_bincount: sequential, in pair order (:205-226)
    return tuple(out)


def compute_xi(healpixs):
    """co.py:77-132"""
    nb = num_bins_r_par * num_bins_r_trans
    weights, r_par, r_trans, z = np.zeros(nb), np.zeros(nb), np.zeros(nb), np.zeros(nb)
    num_pairs = np.zeros(nb, dtype=np.int64)
    for healpix in healpixs:
        for obj1 in objs[healpix]:
            _host.progress(_THIS)
            if obj1.neighbours.size == 0:
                continue
            ang = _host.angle_between_many(obj1, _Rows(obj1.neighbours))  # list branch, :110
            z2 = np.array([obj2.z_qso for obj2 in obj1.neighbours])
            r_comov2 = np.array([obj2.r_comov for obj2 in obj1.neighbours])
            dist_m2 = np.array([obj2.dist_m for obj2 in obj1.neighbours])
            weights2 = np.array([obj2.weights for obj2 in obj1.neighbours])
            res = compute_xi_forest_pairs(obj1.z_qso, obj1.r_comov, obj1.dist_m, obj1.weights, z2,
                                          r_comov2, dist_m2, weights2, ang)
            for target, rebin in zip((weights, r_par, r_trans, z, num_pairs), res):
                target[:len(rebin)] += rebin
            setattr(obj1, "neighbours", None)
    w = weights > 0.
    r_par[w] /= weights[w]
    r_trans[w] /= weights[w]
    z[w] /= weights[w]
    return weights, r_par, r_trans, z, num_pairs
