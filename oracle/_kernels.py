"""ctypes binding of ``libpicca_oracle.so`` (the C restatement in ``picca_oracle.c``).
TEST INFRASTRUCTURE ONLY."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

c_dp = ctypes.POINTER(ctypes.c_double)
c_lp = ctypes.POINTER(ctypes.c_int64)
c_ip = ctypes.POINTER(ctypes.c_int32)


class Params(ctypes.Structure):
    """Field-for-field mirror of ``orc_params`` (the picca.cf / picca.xcf module globals)."""
    _fields_ = [
        ("num_bins_r_par", ctypes.c_int32),
        ("num_bins_r_trans", ctypes.c_int32),
        ("num_model_bins_r_par", ctypes.c_int32),
        ("num_model_bins_r_trans", ctypes.c_int32),
        ("r_par_min", ctypes.c_double),
        ("r_par_max", ctypes.c_double),
        ("r_trans_max", ctypes.c_double),
        ("has_z_min_pairs", ctypes.c_int32),
        ("has_z_max_pairs", ctypes.c_int32),
        ("z_min_pairs", ctypes.c_double),
        ("z_max_pairs", ctypes.c_double),
        ("has_zerr_cut", ctypes.c_int32),
        ("x_correlation", ctypes.c_int32),
        ("zerr_cut_deg", ctypes.c_double),
        ("zerr_cut_kms", ctypes.c_double),
        ("rmu_binning", ctypes.c_int32),
        ("ang_correlation", ctypes.c_int32),
        ("remove_same_half_plate_close_pairs", ctypes.c_int32),
        ("redshift_evolution_in_distortion_matrix", ctypes.c_int32),
        ("z_ref", ctypes.c_double),
        ("alpha", ctypes.c_double),
        ("alpha2", ctypes.c_double),
    ]


def params_from_module(mod, cross=False):
    """Snapshot the module globals (read at call time) into a Params struct."""
    g = lambda name, default=None: getattr(mod, name, default)
    p = Params()
    p.num_bins_r_par = int(g("num_bins_r_par"))
    p.num_bins_r_trans = int(g("num_bins_r_trans"))
    p.num_model_bins_r_par = int(g("num_model_bins_r_par") or p.num_bins_r_par)
    p.num_model_bins_r_trans = int(g("num_model_bins_r_trans") or p.num_bins_r_trans)
    p.r_par_min = float(g("r_par_min"))
    p.r_par_max = float(g("r_par_max"))
    p.r_trans_max = float(g("r_trans_max"))
    p.has_z_min_pairs = int(g("z_min_pairs") is not None)
    p.has_z_max_pairs = int(g("z_max_pairs") is not None)
    p.z_min_pairs = float(g("z_min_pairs") or 0.0)
    p.z_max_pairs = float(g("z_max_pairs") or 0.0)
    has_zerr = (g("zerr_cut_deg") is not None) and not cross
    p.has_zerr_cut = int(has_zerr)
    p.zerr_cut_deg = float(g("zerr_cut_deg") or 0.0)
    p.zerr_cut_kms = float(g("zerr_cut_kms") or 0.0)
    p.x_correlation = int(bool(g("x_correlation", False)))
    p.rmu_binning = int(bool(g("rmu_binning", False)))
    p.ang_correlation = int(bool(g("ang_correlation", False)))
    p.remove_same_half_plate_close_pairs = int(bool(g("remove_same_half_plate_close_pairs", False)))
    p.redshift_evolution_in_distortion_matrix = int(
        bool(g("redshift_evolution_in_distortion_matrix", True)))
    p.z_ref = float(g("z_ref") if g("z_ref") is not None else 0.0)
    p.alpha = float(g("alpha") if g("alpha") is not None else 0.0)
    second = g("alpha_obj") if cross else g("alpha2")
    p.alpha2 = float(second if second is not None else 0.0)
    return p


def build(force=False):
    """Compile the C oracle in place (gcc, no FMA contraction)."""
    so = os.path.join(_HERE, "libpicca_oracle.so")
    src = os.path.join(_HERE, "picca_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libpicca_oracle.so"],
                              stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        assert _LIB.orc_sizeof_params() == ctypes.sizeof(Params)
        _LIB.orc_dmat_auto_pair.restype = ctypes.c_int
        _LIB.orc_dmat_cross_forest.restype = ctypes.c_int
    return _LIB


def dp(a):
    assert a.dtype == np.float64 and a.flags.c_contiguous, (a.dtype, a.flags)
    return a.ctypes.data_as(c_dp)


def lp(a):
    assert a.dtype == np.int64 and a.flags.c_contiguous
    return a.ctypes.data_as(c_lp)


def ip(a):
    if a is None:
        return None
    assert a.dtype == np.int32 and a.flags.c_contiguous
    return a.ctypes.data_as(c_ip)


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def xi_auto_pair(p, d1, d2, ang, same_half_plate, out, ang_correlation=False):
    """cf.compute_xi_forest_pairs_fast on two Delta-like objects; ``out`` = the six rebin arrays
    (weights, xi, r_par, r_trans, z, num_pairs) accumulated in place (cf.py:250-387)."""
    if ang_correlation:  # cf.py:186-208: lambda replaces both distances
        rc1 = dm1 = f64(10.0**d1.log_lambda)
        rc2 = dm2 = f64(10.0**d2.log_lambda)
    else:
        rc1, dm1, rc2, dm2 = f64(d1.r_comov), f64(d1.dist_m), f64(d2.r_comov), f64(d2.dist_m)
    z1, w1, de1 = f64(d1.z), f64(d1.weights), f64(d1.delta)
    z2, w2, de2 = f64(d2.z), f64(d2.weights), f64(d2.delta)
    lib().orc_xi_auto_pair(
        ctypes.byref(p), ctypes.c_int64(z1.size), dp(z1), dp(rc1), dp(dm1), dp(w1), dp(de1),
        ctypes.c_double(d1.z_qso), ctypes.c_int64(z2.size), dp(z2), dp(rc2), dp(dm2), dp(w2),
        dp(de2), ctypes.c_double(d2.z_qso), ctypes.c_double(ang), ctypes.c_int32(same_half_plate),
        dp(out[0]), dp(out[1]), dp(out[2]), dp(out[3]), dp(out[4]), lp(out[5]))


def xi_cross_forest(p, d1, z2, rc2, dm2, w2, ang, out, ang_correlation=False):
    """xcf.compute_xi_forest_pairs_fast (xcf.py:223-322)."""
    if ang_correlation:
        rc1 = dm1 = f64(10.0**d1.log_lambda)
    else:
        rc1, dm1 = f64(d1.r_comov), f64(d1.dist_m)
    z1, w1, de1 = f64(d1.z), f64(d1.weights), f64(d1.delta)
    z2, rc2, dm2, w2, ang = f64(z2), f64(rc2), f64(dm2), f64(w2), f64(ang)
    lib().orc_xi_cross_forest(
        ctypes.byref(p), ctypes.c_int64(z1.size), dp(z1), dp(rc1), dp(dm1), dp(w1), dp(de1),
        ctypes.c_int64(z2.size), dp(z2), dp(rc2), dp(dm2), dp(w2), dp(ang),
        dp(out[0]), dp(out[1]), dp(out[2]), dp(out[3]), dp(out[4]), lp(out[5]))


def dmat_auto_pair(p, d1, d2, ang, same_half_plate, weights_dmat, dmat, r_par_eff, r_trans_eff,
                   z_eff, weight_eff):
    """cf.compute_dmat_forest_pairs_fast (cf.py:520-887)."""
    ll1, rc1, dm1, z1, w1 = (f64(d1.log_lambda), f64(d1.r_comov), f64(d1.dist_m), f64(d1.z),
                             f64(d1.weights))
    ll2, rc2, dm2, z2, w2 = (f64(d2.log_lambda), f64(d2.r_comov), f64(d2.dist_m), f64(d2.z),
                             f64(d2.weights))
    status = lib().orc_dmat_auto_pair(
        ctypes.byref(p), ctypes.c_int64(z1.size), dp(ll1), dp(rc1), dp(dm1), dp(z1), dp(w1),
        ctypes.c_double(d1.z_qso), ctypes.c_int32(int(d1.order)), ctypes.c_int64(z2.size),
        dp(ll2), dp(rc2), dp(dm2), dp(z2), dp(w2), ctypes.c_double(d2.z_qso),
        ctypes.c_int32(int(d2.order)), ctypes.c_double(ang), ctypes.c_int32(same_half_plate),
        dp(weights_dmat), dp(dmat), dp(r_par_eff), dp(r_trans_eff), dp(z_eff), dp(weight_eff))
    if status != 0:
        raise IndexError("negative bin index")  # cf.py:855-856


def dmat_cross_forest(p, d1, rc2, dm2, z2, w2, ang, weights_dmat, dmat, r_par_eff, r_trans_eff,
                      z_eff, weight_eff):
    """xcf.compute_dmat_forest_pairs_fast (xcf.py:427-674)."""
    ll1, rc1, dm1, z1, w1 = (f64(d1.log_lambda), f64(d1.r_comov), f64(d1.dist_m), f64(d1.z),
                             f64(d1.weights))
    rc2, dm2, z2, w2, ang = f64(rc2), f64(dm2), f64(z2), f64(w2), f64(ang)
    status = lib().orc_dmat_cross_forest(
        ctypes.byref(p), ctypes.c_int64(z1.size), dp(ll1), dp(rc1), dp(dm1), dp(z1), dp(w1),
        ctypes.c_int32(int(d1.order)), ctypes.c_int64(z2.size), dp(rc2), dp(dm2), dp(z2), dp(w2),
        dp(ang), dp(weights_dmat), dp(dmat), dp(r_par_eff), dp(r_trans_eff), dp(z_eff),
        dp(weight_eff))
    if status != 0:
        raise IndexError("negative bin index")  # xcf.py:648-649
