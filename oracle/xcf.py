"""Oracle double of the reference module ``picca.xcf`` (hot-path subset).
TEST INFRASTRUCTURE ONLY -- the referee for picca_b200.xcf, never the product.

Restates: fill_neighs (xcf.py:71-123), compute_xi (xcf.py:126-220), compute_dmat (xcf.py:325-424).
"""
import sys

import numpy as np

from . import _host, _kernels

# ---- module globals, names and defaults as reference xcf.py:27-62
num_bins_r_par = None
num_bins_r_trans = None
num_model_bins_r_par = None
num_model_bins_r_trans = None
r_par_max = None
r_par_min = None
r_trans_max = None
z_min_pairs = None
z_max_pairs = None
ang_max = None
nside = None
zerr_cut_deg = None
zerr_cut_kms = None
counter = None
num_data = None
z_ref = None
alpha = None
alpha_obj = None
lambda_abs = None
alpha_abs = None
data = None
objs = None
reject = None
lock = None
cosmo = None
rmu_binning = False
ang_correlation = False
redshift_evolution_in_distortion_matrix = True

_THIS = sys.modules[__name__]


class _ObjCatalogue:
    def __init__(self, cat):
        self.z_qso = np.array([o.z_qso for o in cat.objs], dtype=np.float64)
        self.has_dist = all(o.r_comov is not None for o in cat.objs)
        if self.has_dist:
            self.r_comov = np.array([o.r_comov for o in cat.objs], dtype=np.float64)
            self.dist_m = np.array([o.dist_m for o in cat.objs], dtype=np.float64)
        self.weights = np.array([o.weights for o in cat.objs], dtype=np.float64)


def _obj_arrays():
    cat = _host.catalogue(objs)
    if not hasattr(cat, "_obj"):
        cat._obj = _ObjCatalogue(cat)
    return cat, cat._obj


def fill_neighs(healpixs):
    """xcf.py:71-123; neighbours stored as an index array into the ascending-healpix object
    catalogue plus the object list itself (``delta.neighbours`` is a NumPy object array)."""
    cat, arr = _obj_arrays()
    for healpix in healpixs:
        for delta in data[healpix]:
            ang = _host.angle_between_many(delta, cat)
            w = (cat.thingid != delta.thingid) & (ang < ang_max)
            if zerr_cut_deg is not None:  # xcf.py:102-115
                ang_deg = 180.0 / np.pi * ang
                z_qq = 0.5 * (delta.z_qso + arr.z_qso)
                dv_kms = np.abs(delta.z_qso - arr.z_qso) / (1 + z_qq)
                dv_kms *= _host.SPEED_LIGHT
                zerr_cut_mask = ang_deg < zerr_cut_deg
                zerr_cut_mask &= dv_kms < zerr_cut_kms
                w &= ~zerr_cut_mask
            if not ang_correlation:  # xcf.py:117-121
                f = r_trans_max if rmu_binning else 1
                w &= (delta.r_comov[0] - arr.r_comov) * np.cos(ang / 2.) < r_par_max * f
                w &= (delta.r_comov[-1] - arr.r_comov) * np.cos(ang / 2.) > r_par_min * f
            idx = np.nonzero(w)[0]
            neighbours = np.empty(idx.size, dtype=object)
            for k, q in enumerate(idx):
                neighbours[k] = cat.objs[q]
            delta.neighbours = neighbours
            delta._neighbour_index = idx


def compute_xi(healpixs):
    """xcf.py:126-220."""
    p = _kernels.params_from_module(_THIS, cross=True)
    cat, arr = _obj_arrays()
    nb = num_bins_r_par * num_bins_r_trans
    out = [np.zeros(nb) for _ in range(5)] + [np.zeros(nb, dtype=np.int64)]
    for healpix in healpixs:
        for delta in data[healpix]:
            _host.progress(_THIS)
            if delta.neighbours.size != 0:
                idx = delta._neighbour_index
                ang = _host.angle_between_many(delta, cat, idx)
                z_qso = arr.z_qso[idx]
                weights_qso = arr.weights[idx]
                if ang_correlation:
                    lambda_qso = np.array([10.0**obj.log_lambda for obj in delta.neighbours])
                    _kernels.xi_cross_forest(p, delta, z_qso, lambda_qso, lambda_qso, weights_qso,
                                             ang, out, ang_correlation=True)
                else:
                    _kernels.xi_cross_forest(p, delta, z_qso, arr.r_comov[idx], arr.dist_m[idx],
                                             weights_qso, ang, out)
            setattr(delta, "neighbours", None)
    weights, xi, r_par, r_trans, z, num_pairs = out
    w = weights > 0
    xi[w] /= weights[w]
    r_par[w] /= weights[w]
    r_trans[w] /= weights[w]
    z[w] /= weights[w]
    return weights, xi, r_par, r_trans, z, num_pairs


def compute_dmat(healpixs):
    """xcf.py:325-424."""
    p = _kernels.params_from_module(_THIS, cross=True)
    cat, arr = _obj_arrays()
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
            w = np.random.rand(len(delta1.neighbours)) > reject  # xcf.py:379
            if w.sum() == 0:
                continue  # xcf.py:380-381 (Q7: forest not counted, neighbours not cleared)
            num_pairs += len(delta1.neighbours)
            num_pairs_used += w.sum()
            idx = delta1._neighbour_index[w]
            ang = _host.angle_between_many(delta1, cat, idx)
            _kernels.dmat_cross_forest(p, delta1, arr.r_comov[idx], arr.dist_m[idx],
                                       arr.z_qso[idx], arr.weights[idx], ang, weights_dmat, dmat,
                                       r_par_eff, r_trans_eff, z_eff, weight_eff)
            setattr(delta1, "neighbours", None)
    dmat = dmat.reshape(nb, nbm)
    return (weights_dmat, dmat, r_par_eff, r_trans_eff, z_eff, weight_eff, num_pairs,
            num_pairs_used)


# rest wavelengths of the transitions the tests use (reference constants.py:243-300)
absorber_igm = {"LYA": 1215.67, "SiIII(1207)": 1206.500, "SiII(1190)": 1190.4158,
                "SiII(1193)": 1193.2897, "SiII(1260)": 1260.4221, "CIV(eff)": 1549.06}


def compute_metal_dmat(healpixs, abs_igm="SiII(1526)"):
    """xcf.py:677-835."""
    nb = num_bins_r_par * num_bins_r_trans
    nbm = num_model_bins_r_par * num_model_bins_r_trans
    dmat = np.zeros(nb * nbm)
    weights_dmat = np.zeros(nb)
    r_par_eff, r_trans_eff, z_eff, weight_eff = (np.zeros(nbm) for _ in range(4))
    num_pairs = 0
    num_pairs_used = 0
    for healpix in healpixs:
        for delta1 in data[healpix]:
            _host.progress(_THIS)
            z1_abs1 = 10**delta1.log_lambda / absorber_igm[abs_igm] - 1        # :729
            r_comov1_abs1 = cosmo.get_r_comov(z1_abs1)
            dist_m1_abs1 = cosmo.get_dist_m(z1_abs1)
            w = z1_abs1 < delta1.z_qso                                         # :735
            r_comov1, dist_m1, weights1 = delta1.r_comov[w], delta1.dist_m[w], delta1.weights[w]
            z1_abs1, r_comov1_abs1, dist_m1_abs1 = z1_abs1[w], r_comov1_abs1[w], dist_m1_abs1[w]
            if r_comov1.size == 0:                                             # :741-742
                continue
            w = np.random.rand(len(delta1.neighbours)) > reject               # :744
            num_pairs += len(delta1.neighbours)
            num_pairs_used += w.sum()
            for obj2 in [o for o, keep in zip(delta1.neighbours, w) if keep]:
                ang = _host.angle_between_one(delta1, obj2)
                r_par = (r_comov1 - obj2.r_comov) * np.cos(ang / 2)           # :755-757
                r_trans = (dist_m1 + obj2.dist_m) * np.sin(ang / 2)
                weights12 = weights1 * obj2.weights
                w = (r_par > r_par_min) & (r_par < r_par_max) & (r_trans < r_trans_max)
                bins_r_par = ((r_par - r_par_min) / (r_par_max - r_par_min) *
                              num_bins_r_par).astype(int)
                bins_r_trans = (r_trans / r_trans_max * num_bins_r_trans).astype(int)
                bins = bins_r_trans + num_bins_r_trans * bins_r_par
                rebin = np.bincount(bins[w], weights=weights12[w])
                weights_dmat[:len(rebin)] += rebin
                r_par_abs = (r_comov1_abs1 - obj2.r_comov) * np.cos(ang / 2)  # :767-768
                r_trans_abs = (dist_m1_abs1 + obj2.dist_m) * np.sin(ang / 2)
                z_weight_evol = ((1.0 + z1_abs1) / (1.0 + z_ref))**(alpha_abs[abs_igm] - 1.0)
                model_bins_r_par = ((r_par_abs - r_par_min) / (r_par_max - r_par_min) *
                                    num_model_bins_r_par).astype(int)
                model_bins_r_trans = (r_trans_abs / r_trans_max *
                                      num_model_bins_r_trans).astype(int)
                model_bins = model_bins_r_trans + num_model_bins_r_trans * model_bins_r_par
                w &= (r_par_abs > r_par_min) & (r_par_abs < r_par_max) & \
                    (r_trans_abs < r_trans_max)                                 # :785-789
                for target, index, values in (
                        (dmat, model_bins[w] + nbm * bins[w], weights12[w] * z_weight_evol[w]),
                        (r_par_eff, model_bins[w], r_par_abs[w] * weights12[w] * z_weight_evol[w]),
                        (r_trans_eff, model_bins[w],
                         r_trans_abs[w] * weights12[w] * z_weight_evol[w]),
                        (z_eff, model_bins[w],
                         (z1_abs1 + obj2.z_qso)[w] / 2 * weights12[w] * z_weight_evol[w]),
                        (weight_eff, model_bins[w], weights12[w] * z_weight_evol[w])):
                    rebin = np.bincount(index, weights=values)
                    target[:len(rebin)] += rebin
            setattr(delta1, "neighbours", None)
    return (weights_dmat, dmat.reshape(nb, nbm), r_par_eff, r_trans_eff, z_eff, weight_eff,
            num_pairs, num_pairs_used)


# ---- Wick expansion, terms T1-T4 (xcf.py:838-1153, :1219-1351).  Globals as the reference: the
# script fills get_variance_1d / xi_1d per delta.fname (picca_xwick.py:394-409).
get_variance_1d = {}
xi_1d = {}
max_diagram = None
xi_wick = None


def compute_wickT1234_pairs(ang, r_comov1, r_comov2, z1, z2, weights1, weights2, weighted_xi_1d_1,
                            weights_wick, num_pairs_wick, t1, t2, t3, t4):
    """xcf.py:1219-1351: the C restatement ``orc_wick_t1234_forest``, accumulating in place."""
    import ctypes
    lib = _kernels.lib()
    p = _kernels.params_from_module(_THIS, cross=True)
    f64, dp, lp = _kernels.f64, _kernels.dp, _kernels.lp
    a, r1, r2, zz1, zz2, w1, w2 = (f64(v) for v in (ang, r_comov1, r_comov2, z1, z2, weights1,
                                                    weights2))
    x1 = f64(weighted_xi_1d_1)
    for arr in (weights_wick, num_pairs_wick, t1, t2, t3, t4):
        assert arr.flags.c_contiguous
    lib.orc_wick_t1234_forest(
        ctypes.byref(p), ctypes.c_int64(len(r1)), dp(r1), dp(zz1), dp(w1), dp(x1),
        ctypes.c_int64(len(r2)), dp(a), dp(r2), dp(zz2), dp(w2), dp(weights_wick),
        lp(num_pairs_wick), dp(t1), dp(t2), dp(t3), dp(t4))


def compute_wick_terms(healpixs):
    """xcf.py:838-1153 for the diagrams T1-T4 (``xi_wick is None or max_diagram <= 4``)."""
    if xi_wick is not None and max_diagram is not None and max_diagram > 4:
        raise NotImplementedError("oracle: Wick diagrams T5-T6 are not restated")
    nb = num_bins_r_par * num_bins_r_trans
    t1, t2, t3, t4, t5, t6 = (np.zeros((nb, nb)) for _ in range(6))
    weights_wick = np.zeros(nb)
    num_pairs_wick = np.zeros(nb, dtype=np.int64)
    num_pairs = 0
    num_pairs_used = 0
    for healpix in healpixs:
        num_pairs += len(data[healpix])
        w = np.random.rand(len(data[healpix])) > reject      # xcf.py:889
        num_pairs_used += w.sum()
        if w.sum() == 0:
            continue
        for delta1 in [delta for index, delta in enumerate(data[healpix]) if w[index]]:
            _host.progress(_THIS)
            if delta1.neighbours.size == 0:
                continue
            variance_1d = get_variance_1d[delta1.fname](delta1.log_lambda)
            weights1 = delta1.weights
            weighted_xi_1d_1 = ((weights1 * weights1[:, None]) *
                                xi_1d[delta1.fname](abs(delta1.log_lambda -
                                                        delta1.log_lambda[:, None])) *
                                np.sqrt(variance_1d * variance_1d[:, None]))   # xcf.py:907-914
            neighbours = delta1.neighbours
            ang12 = np.array([_host.angle_between_one(delta1, obj2) for obj2 in neighbours])
            r_comov2 = np.array([obj2.r_comov for obj2 in neighbours])
            z2 = np.array([obj2.z_qso for obj2 in neighbours])
            weights2 = np.array([obj2.weights for obj2 in neighbours])
            compute_wickT1234_pairs(ang12, delta1.r_comov, r_comov2, delta1.z, z2, weights1,
                                    weights2, weighted_xi_1d_1, weights_wick, num_pairs_wick,
                                    t1, t2, t3, t4)
    return weights_wick, num_pairs_wick, num_pairs, num_pairs_used, t1, t2, t3, t4, t5, t6
