"""Oracle double of the reference module ``picca.cf`` (hot-path subset): same globals, same
functions, same return tuples, computed on the CPU by the C restatement.
TEST INFRASTRUCTURE ONLY -- the referee for picca_b200.cf, never the product.

Restates: fill_neighs (cf.py:82-135), compute_xi (cf.py:138-247), compute_dmat (cf.py:390-517),
compute_metal_dmat (cf.py:890-1232), compute_wick_terms for max_diagram <= 3 with
compute_wickT123_pairs (cf.py:1326-1494, :1497-1626; prepared for the next round: no CUDA
counterpart yet).
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


# rest wavelengths of the transitions the tests use (reference constants.py:243-300); the live
# reference's table is used by the pinning tests
absorber_igm = {"LYA": 1215.67, "SiIII(1207)": 1206.500, "SiII(1190)": 1190.4158,
                "SiII(1193)": 1193.2897, "SiII(1260)": 1260.4221, "CIV(eff)": 1549.06}


def _metal_side(delta, name):
    """cf.py:941-960 (and :973-992, :1092-1125 with the roles swapped): the pixels of one forest
    that are consistent with the quasar redshift for absorber ``name``."""
    z_abs = 10**delta.log_lambda / absorber_igm[name] - 1
    r_comov_abs = cosmo.get_r_comov(z_abs)
    dist_m_abs = cosmo.get_dist_m(z_abs)
    w = z_abs < delta.z_qso
    return (delta.r_comov[w], delta.dist_m[w], delta.weights[w], r_comov_abs[w], dist_m_abs[w],
            z_abs[w])


def _metal_pass(delta1, delta2, name1, name2, ang, same_half_plate, out):
    """One (absorber of forest 1, absorber of forest 2) pass: cf.py:994-1087, repeated with the
    absorbers swapped at cf.py:1126-1217."""
    weights_dmat, dmat, r_par_eff, r_trans_eff, z_eff, weight_eff = out
    r_comov1, dist_m1, weights1, r_comov1_abs, dist_m1_abs, z1_abs = _metal_side(delta1, name1)
    r_comov2, dist_m2, weights2, r_comov2_abs, dist_m2_abs, z2_abs = _metal_side(delta2, name2)
    r_par = (r_comov1[:, None] - r_comov2) * np.cos(ang / 2)
    if not x_correlation:
        r_par = abs(r_par)
    r_trans = (dist_m1[:, None] + dist_m2) * np.sin(ang / 2)
    weights12 = weights1[:, None] * weights2
    bins_r_par = np.floor((r_par - r_par_min) / (r_par_max - r_par_min) *
                          num_bins_r_par).astype(int)
    bins_r_trans = (r_trans / r_trans_max * num_bins_r_trans).astype(int)
    if remove_same_half_plate_close_pairs and same_half_plate:
        weights12[abs(r_par) < (r_par_max - r_par_min) / num_bins_r_par] = 0.0
    bins = bins_r_trans + num_bins_r_trans * bins_r_par
    w = (bins_r_par < num_bins_r_par) & (bins_r_trans < num_bins_r_trans) & (bins_r_par >= 0)
    rebin = np.bincount(bins[w], weights=weights12[w])
    weights_dmat[:len(rebin)] += rebin

    r_par_m = (r_comov1_abs[:, None] - r_comov2_abs) * np.cos(ang / 2)
    if not x_correlation:
        r_par_m = abs(r_par_m)
    r_trans_m = (dist_m1_abs[:, None] + dist_m2_abs) * np.sin(ang / 2)
    z_weight_evol = ((1 + z1_abs[:, None])**(alpha_abs[name1] - 1) *
                     (1 + z2_abs)**(alpha_abs[name2] - 1) /
                     (1 + z_ref)**(alpha_abs[name1] + alpha_abs[name2] - 2))
    model_bins_r_par = np.floor((r_par_m - r_par_min) / (r_par_max - r_par_min) *
                                num_model_bins_r_par).astype(int)
    model_bins_r_trans = (r_trans_m / r_trans_max * num_model_bins_r_trans).astype(int)
    model_bins = model_bins_r_trans + num_model_bins_r_trans * model_bins_r_par
    w &= ((model_bins_r_par < num_model_bins_r_par) &
          (model_bins_r_trans < num_model_bins_r_trans) & (model_bins_r_par >= 0))
    nbm = num_model_bins_r_par * num_model_bins_r_trans
    for target, index, values in (
            (dmat, model_bins[w] + nbm * bins[w], weights12[w] * z_weight_evol[w]),
            (r_par_eff, model_bins[w], r_par_m[w] * weights12[w] * z_weight_evol[w]),
            (r_trans_eff, model_bins[w], r_trans_m[w] * weights12[w] * z_weight_evol[w]),
            (z_eff, model_bins[w],
             (z1_abs[:, None] + z2_abs)[w] / 2 * weights12[w] * z_weight_evol[w]),
            (weight_eff, model_bins[w], weights12[w] * z_weight_evol[w])):
        rebin = np.bincount(index, weights=values)
        target[:len(rebin)] += rebin


def compute_metal_dmat(healpixs, abs_igm1="LYA", abs_igm2="SiIII(1207)"):
    """cf.py:890-1232."""
    nb = num_bins_r_par * num_bins_r_trans
    nbm = num_model_bins_r_par * num_model_bins_r_trans
    out = (np.zeros(nb), np.zeros(nb * nbm), np.zeros(nbm), np.zeros(nbm), np.zeros(nbm),
           np.zeros(nbm))
    num_pairs = 0
    num_pairs_used = 0
    for healpix in healpixs:
        for delta1 in data[healpix]:
            _host.progress(_THIS)
            w = np.random.rand(len(delta1.neighbours)) > reject  # cf.py:943
            num_pairs += len(delta1.neighbours)
            num_pairs_used += w.sum()
            for delta2 in [d for d, keep in zip(delta1.neighbours, w) if keep]:
                shp = _host.same_half_plate(delta1, delta2) \
                    if remove_same_half_plate_close_pairs else False
                ang = _host.angle_between_one(delta1, delta2)
                _metal_pass(delta1, delta2, abs_igm1, abs_igm2, ang, shp, out)
                if ((not x_correlation) and (abs_igm1 != abs_igm2)) or \
                        (x_correlation and (lambda_abs == lambda_abs2)):  # cf.py:1089-1091
                    _metal_pass(delta1, delta2, abs_igm2, abs_igm1, ang, shp, out)
            setattr(delta1, "neighbours", None)
    weights_dmat, dmat, r_par_eff, r_trans_eff, z_eff, weight_eff = out
    return (weights_dmat, dmat.reshape(nb, nbm), r_par_eff, r_trans_eff, z_eff, weight_eff,
            num_pairs, num_pairs_used)


# ---- Wick expansion, terms T1-T3 (cf.py:1326-1626).  Globals as the reference: the script fills
# get_variance_1d / xi_1d per delta.fname and sets max_diagram (picca_wick.py:336, :393-416).
get_variance_1d = {}
xi_1d = {}
max_diagram = None


def _weighted_xi_1d(delta):
    """cf.py:1413-1421"""
    variance_1d = get_variance_1d[delta.fname](delta.log_lambda)
    weights = delta.weights
    return ((weights * weights[:, None]) *
            xi_1d[delta.fname](abs(delta.log_lambda - delta.log_lambda[:, None])) *
            np.sqrt(variance_1d * variance_1d[:, None]))


def compute_wickT123_pairs(r_comov1, r_comov2, ang, weights1, weights2, z1, z2, weighted_xi_1d_1,
                           weighted_xi_1d_2, weights_wick, num_pairs_wick, t1, t2, t3):
    """cf.py:1497-1626: the C restatement ``orc_wick_t123_pair`` (statement by statement: the
    selected pixel pairs in the reference's order, then its double loop over pairs of list
    entries), accumulating in place like the Numba function."""
    import ctypes
    lib = _kernels.lib()
    p = _kernels.params_from_module(_THIS)
    f64, dp, lp = _kernels.f64, _kernels.dp, _kernels.lp
    r1, r2, w1, w2, zz1, zz2 = (f64(a) for a in (r_comov1, r_comov2, weights1, weights2, z1, z2))
    x1, x2 = f64(weighted_xi_1d_1), f64(weighted_xi_1d_2)
    for arr in (weights_wick, num_pairs_wick, t1, t2, t3):
        assert arr.flags.c_contiguous
    lib.orc_wick_t123_pair(
        ctypes.byref(p), ctypes.c_int64(len(r1)), dp(r1), ctypes.c_int64(len(r2)), dp(r2),
        ctypes.c_double(float(ang)), dp(w1), dp(w2), dp(zz1), dp(zz2), dp(x1), dp(x2),
        dp(weights_wick), lp(num_pairs_wick), dp(t1), dp(t2), dp(t3))


def compute_wick_terms(healpixs):
    """cf.py:1326-1494 for max_diagram <= 3 (T4-T6 stay zero)."""
    if max_diagram is not None and max_diagram > 3:
        raise NotImplementedError("oracle: Wick diagrams T4-T6 are not restated")
    nb = num_bins_r_par * num_bins_r_trans
    t1, t2, t3, t4, t5, t6 = (np.zeros((nb, nb)) for _ in range(6))
    weights_wick = np.zeros(nb)
    num_pairs_wick = np.zeros(nb, dtype=np.int64)
    num_pairs = 0
    num_pairs_used = 0
    for healpix in healpixs:
        w = np.random.rand(len(data[healpix])) > reject      # :1378, one number per FOREST
        num_pairs += len(data[healpix])
        num_pairs_used += w.sum()
        if w.sum() == 0:
            continue
        for delta1 in [delta for index, delta in enumerate(data[healpix]) if w[index]]:
            _host.progress(_THIS)
            if len(delta1.neighbours) == 0:
                continue
            weighted_xi_1d_1 = _weighted_xi_1d(delta1)
            for delta2 in delta1.neighbours:
                ang12 = _host.angle_between_one(delta1, delta2)
                compute_wickT123_pairs(delta1.r_comov, delta2.r_comov, ang12, delta1.weights,
                                       delta2.weights, delta1.z, delta2.z, weighted_xi_1d_1,
                                       _weighted_xi_1d(delta2), weights_wick, num_pairs_wick,
                                       t1, t2, t3)
    return weights_wick, num_pairs_wick, num_pairs, num_pairs_used, t1, t2, t3, t4, t5, t6
