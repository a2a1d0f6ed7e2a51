"""Oracle double of the reference module ``picca.cf`` (hot-path subset): same globals, same
functions, same return tuples, computed on the CPU by the C restatement.
TEST INFRASTRUCTURE ONLY -- the referee for picca_b200.cf, never the product.

Restates: fill_neighs (cf.py:82-135), compute_xi (cf.py:138-247), compute_dmat (cf.py:390-517).
"""
import sys

import numpy as np

from . import _host, _kernels

# ---- module globals, names and defaults as reference cf.py:28-67
num_bins_r_par = None
num_bins_r_trans = None
num_model_bins_r_trans = None
num_model_bins_r_par = None
r_par_max = None
r_par_min = None
z_min_pairs = None
z_max_pairs = None
r_trans_max = None
ang_max = None
nside = None
zerr_cut_deg = None
zerr_cut_kms = None
counter = None
num_data = None
num_data2 = None
z_ref = None
alpha = None
alpha2 = None
alpha_abs = None
lambda_abs = None
lambda_abs2 = None
data = None
data2 = None
cosmo = None
reject = None
lock = None
x_correlation = False
rmu_binning = False
ang_correlation = False
remove_same_half_plate_close_pairs = False
redshift_evolution_in_distortion_matrix = True

_THIS = sys.modules[__name__]


def fill_neighs(healpixs):
    """cf.py:82-135.  Candidates: every forest of the other catalogue in ascending-healpix, list
    order (a superset of query_disc's pixels; the exact ``ang < ang_max`` filter follows)."""
    other = _host.catalogue(data2 if data2 is not None else data)
    for healpix in healpixs:
        for delta in data[healpix]:
            ang = _host.angle_between_many(delta, other)
            w = (other.thingid != delta.thingid) & (ang < ang_max)
            if data2 is None:
                w &= delta.ra > other.ra  # cf.py:129-135
            delta.neighbours = [other.objs[k] for k in np.nonzero(w)[0]]


def compute_xi(healpixs):
    """cf.py:138-247."""
    p = _kernels.params_from_module(_THIS)
    nb = num_bins_r_par * num_bins_r_trans
    out = [np.zeros(nb) for _ in range(5)] + [np.zeros(nb, dtype=np.int64)]
    for healpix in healpixs:
        for delta1 in data[healpix]:
            _host.progress(_THIS)
            for delta2 in delta1.neighbours:
                ang = _host.angle_between_one(delta1, delta2)
                if remove_same_half_plate_close_pairs:
                    shp = _host.same_half_plate(delta1, delta2)
                else:
                    shp = False
                _kernels.xi_auto_pair(p, delta1, delta2, float(ang), int(shp), out,
                                      ang_correlation=ang_correlation)
            setattr(delta1, "neighbours", None)
    weights, xi, r_par, r_trans, z, num_pairs = out
    w = weights > 0
    xi[w] /= weights[w]
    r_par[w] /= weights[w]
    r_trans[w] /= weights[w]
    z[w] /= weights[w]
    return weights, xi, r_par, r_trans, z, num_pairs


def compute_dmat(healpixs):
    """cf.py:390-517 (consumes the global legacy NumPy RNG exactly like the reference)."""
    p = _kernels.params_from_module(_THIS)
    nb = num_bins_r_par * num_bins_r_trans
    nbm = num_model_bins_r_par * num_model_bins_r_trans
    dmat = np.zeros(nb * nbm)
    weights_dmat = np.zeros(nb)
    r_par_eff = np.zeros(nbm)
    r_trans_eff = np.zeros(nbm)
    z_eff = np.zeros(nbm)
    weight_eff = np.zeros(nbm)
    num_pairs = 0
    num_pairs_used = 0
    for healpix in healpixs:
        for delta1 in data[healpix]:
            _host.progress(_THIS)
            if delta1.order is None:
                raise RuntimeError("Trying to compute the distortion matrix but "
                                   "order is not defined for the deltas. "
                                   "Check previous warning to solve this issue")
            w = np.random.rand(len(delta1.neighbours)) > reject  # cf.py:444
            num_pairs += len(delta1.neighbours)
            num_pairs_used += w.sum()
            for delta2 in [d for d, keep in zip(delta1.neighbours, w) if keep]:
                if remove_same_half_plate_close_pairs:
                    shp = _host.same_half_plate(delta1, delta2)
                else:
                    shp = False
                if delta2.order is None:
                    raise RuntimeError("Trying to compute the distortion matrix but "
                                       "order is not defined for the deltas. "
                                       "Check previous warning to solve this issue")
                ang = _host.angle_between_one(delta1, delta2)
                _kernels.dmat_auto_pair(p, delta1, delta2, float(ang), int(shp), weights_dmat,
                                        dmat, r_par_eff, r_trans_eff, z_eff, weight_eff)
            setattr(delta1, "neighbours", None)
    dmat = dmat.reshape(nb, nbm)
    return (weights_dmat, dmat, r_par_eff, r_trans_eff, z_eff, weight_eff, num_pairs,
            num_pairs_used)
