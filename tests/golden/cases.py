"""Definition of the golden cases: seeded synthetic inputs (regenerated identically wherever the
tests run) + the module configuration.  ``make_golden.py`` runs the LIVE reference on them and
stores its outputs in ``golden_*.npz``; the tests replay the same cases on the oracle (CPU) and on
the CUDA path (GPU)."""
import numpy as np

from picca_b200 import synth
from tests import helpers

R60 = dict(r_par_max=60., r_trans_max=60., num_bins_r_par=15, num_bins_r_trans=15,
           num_model_bins_r_par=15, num_model_bins_r_trans=15)

# angular correlation: "r_par" = wavelength ratio, "r_trans" = angle (rad), ang_max = r_trans_max
ANGL = dict(ang_correlation=True, r_par_min=1., r_par_max=1.1, r_trans_max=0.02,
            num_bins_r_par=20, num_bins_r_trans=10)

CF_CASES = {
    "default": dict(R60),
    "half_plate": dict(R60, remove_same_half_plate_close_pairs=True),
    # the synthetic forests have one plate each: `plates` deals them onto 7 shared plates so that
    # the same-half-plate rule (cf.py:171-183, :378-380) actually removes pairs
    "half_plate_shared": dict(R60, remove_same_half_plate_close_pairs=True, plates=True),
    "zcuts": dict(R60, z_min_pairs=2.0, z_max_pairs=2.6),
    "zerr": dict(R60, zerr_cut_deg=0.5, zerr_cut_kms=40000.),
    "rmu": dict(R60, rmu_binning=True, r_par_min=0., r_par_max=1.),
    "prod": dict(r_par_max=200., r_trans_max=200., num_bins_r_par=50, num_bins_r_trans=50),
    "cross": dict(R60, x_correlation=True, r_par_min=-60., num_bins_r_par=30, second=True),
    # picca_cf_angl.py (:212-222): wavelength ratio x angle, cf.py:186-208, :350-354
    "angl": dict(ANGL),
    "angl_half_plate": dict(ANGL, remove_same_half_plate_close_pairs=True, plates=True),
    "angl_cross": dict(ANGL, x_correlation=True, r_par_min=0.92, r_par_max=1.08, second=True),
}

DMAT_CASES = {
    "default": dict(R60, reject=0.9),
    "noevol_halfplate": dict(R60, reject=0.9, redshift_evolution_in_distortion_matrix=False,
                             remove_same_half_plate_close_pairs=True),
    "zcuts": dict(R60, reject=0.9, z_min_pairs=2.0, z_max_pairs=2.6),
    "coef2": dict(R60, reject=0.95, num_model_bins_r_par=30, num_model_bins_r_trans=30),
    "cross": dict(R60, reject=0.9, x_correlation=True, r_par_min=-60., num_bins_r_par=30,
                  num_model_bins_r_par=30, second=True),
}

ALPHA_ABS = {"LYA": 2.9, "SiIII(1207)": 1., "SiII(1190)": 1., "CIV(eff)": 1., "TEST(1045)": 1.7}
# a made-up transition just redward of the blue end of the synthetic forests: only ~20 % of the
# forests have a pixel with z_abs < z_qso, the others are skipped before the --rej draw
# (xcf.py:741-742)
EXTRA_ABSORBERS = {"TEST(1045)": 1045.05}
# (abs_igm1, abs_igm2) + module configuration of cf.compute_metal_dmat (cf.py:890-1232)
METAL_CASES = {
    "lya_si3": dict(R60, reject=0.9, pair=("LYA", "SiIII(1207)")),          # both passes
    "si2_si2": dict(R60, reject=0.9, pair=("SiII(1190)", "SiII(1190)")),    # z_abs < z_qso filter
    "si3_si2_coef2": dict(R60, reject=0.9, pair=("SiIII(1207)", "SiII(1190)"),
                          num_model_bins_r_par=30, num_model_bins_r_trans=30),
    "civ_far": dict(R60, reject=0.9, pair=("LYA", "CIV(eff)")),  # model bins all out of range
    "cross": dict(R60, reject=0.9, pair=("LYA", "SiII(1190)"), x_correlation=True, r_par_min=-60.,
                  num_bins_r_par=30, num_model_bins_r_par=30, second=True,
                  remove_same_half_plate_close_pairs=True, lambda_abs="LYA", lambda_abs2="LYA"),
}

XCF_BASE = dict(r_par_max=60., r_par_min=-60., r_trans_max=60., num_bins_r_par=30,
                num_bins_r_trans=15, num_model_bins_r_par=30, num_model_bins_r_trans=15,
                alpha_obj=1.44)
XCF_CASES = {
    "default": dict(XCF_BASE),
    "zcuts": dict(XCF_BASE, z_min_pairs=2.0, z_max_pairs=2.6),
    "zerr": dict(XCF_BASE, zerr_cut_deg=0.5, zerr_cut_kms=40000.),
    "rmu": dict(XCF_BASE, rmu_binning=True, r_par_min=-1., r_par_max=1.),
    # picca_xcf_angl.py (:268-276): xcf.py:161-182, :293-295 (no r_par pre-filter, :117)
    "angl": dict(ANGL, r_par_min=0.9, alpha_obj=1.44),
    "angl_zcuts": dict(ANGL, r_par_min=0.9, alpha_obj=1.44, z_min_pairs=2.0, z_max_pairs=2.6),
}
XMETAL_CASES = {
    "si2": dict(XCF_BASE, reject=0.8, abs_igm="SiII(1190)"),
    "si3_coef2": dict(XCF_BASE, reject=0.8, abs_igm="SiIII(1207)", num_model_bins_r_par=60,
                      num_model_bins_r_trans=30),
    "edge_skip": dict(XCF_BASE, reject=0.5, abs_igm="TEST(1045)"),
}
CO_BASE = dict(r_par_max=80., r_par_min=0., r_trans_max=80., num_bins_r_par=20,
               num_bins_r_trans=20, z_cut_min=0., z_cut_max=10.)
CO_CASES = {
    "dd": dict(CO_BASE, type_corr="DD"),
    "zcut": dict(CO_BASE, type_corr="DD", z_cut_min=2.2, z_cut_max=2.8),
    "xdd": dict(CO_BASE, type_corr="xDD", x_correlation=True, r_par_min=-80., num_bins_r_par=40,
                second=True),
    "dr": dict(CO_BASE, type_corr="DR", x_correlation=True, second=True),  # abs although crossed
}
XDMAT_CASES = {
    "default": dict(XCF_BASE, reject=0.8),
    "noevol": dict(XCF_BASE, reject=0.8, redshift_evolution_in_distortion_matrix=False),
    "zcuts": dict(XCF_BASE, reject=0.8, z_min_pairs=2.0, z_max_pairs=2.6),
}


def share_plates(data):
    """Deal the forests onto 7 plates / 1000 fibres (deterministic in the thingid)."""
    for forests_ in data.values():
        for d in forests_:
            d.plate = 1 + int(d.thingid) % 7
            d.fiberid = 1 + (int(d.thingid) * 131) % 1000
    return data


# Wick expansion (cf.compute_wick_terms T1-T3, xcf.compute_wick_terms T1-T4): per-forest --rej draw,
# 1-D inputs as picca_wick.py builds them (scipy interp1d, nearest, extrapolating; :393-417)
WICK_CASES = {
    "default": dict(R60, reject=0.8, max_diagram=3),
    "coarse_x": dict(R60, reject=0.7, max_diagram=3, x_correlation=True, r_par_min=-60.,
                     num_bins_r_par=12, num_bins_r_trans=6, second=True, alpha2=1.7),
}
XWICK_CASES = {
    "default": dict(XCF_BASE, reject=0.5, max_diagram=4),
    "coarse": dict(XCF_BASE, reject=0.3, max_diagram=4, num_bins_r_par=10, num_bins_r_trans=5,
                   alpha_obj=1.),
}


def wick_1d(fname):
    """(get_variance_1d, xi_1d) interpolators of one delta sample (deterministic tables)."""
    from scipy.interpolate import interp1d
    shift = 0. if fname == "D1" else 0.013
    ll = 3.55 + 5e-4 * np.arange(420)
    var = 0.05 + 0.1 * (ll - 3.55) + 0.01 * np.sin(40. * ll) + shift
    dll = 5e-4 * np.arange(260)
    xi = np.exp(-dll / 3e-3) * np.cos(dll / (2e-3 + shift)) + 0.02
    return (interp1d(ll, var, kind="nearest", fill_value="extrapolate"),
            interp1d(dll, xi, kind="nearest", fill_value="extrapolate"))


def set_fname(data, fname):
    for forests_ in data.values():
        for d in forests_:
            d.fname = fname
    return data


def forests(second=False, plates=False):
    """(data, num_data, z_min, cosmo); ``second`` gives the independent second sample used by the
    delta x delta cross-correlation cases; ``plates``: see ``share_plates``."""
    if second:
        data, num, z_min, _, cosmo = helpers.small_sample(n=200, seed=23, max_pix=100,
                                                          id_offset=5000)
    else:
        data, num, z_min, _, cosmo = helpers.small_sample(n=300, seed=11, max_pix=120)
    if plates:
        share_plates(data)
    return data, num, z_min, cosmo


def dmat_forests(second=False, **kw):
    """smaller forests: the as-written reference dmat is O(N_pairs * U) per forest pair."""
    if second:
        data, num, z_min, _, cosmo = helpers.small_sample(n=80, seed=29, max_pix=60, side_deg=3.,
                                                          id_offset=5000, **kw)
    else:
        data, num, z_min, _, cosmo = helpers.small_sample(n=120, seed=17, max_pix=70, side_deg=3.,
                                                          **kw)
    return data, num, z_min, cosmo


def xwick_forests():
    """no zero-weight pixels: xcf.compute_wickT1234_pairs divides by the pixel weight
    (xcf.py:1316, :1330) and Numba raises ZeroDivisionError on 0/0"""
    return dmat_forests(zero_weight_frac=0.)


def quasars(cosmo):
    objs, z_min = synth.make_quasars(400, seed=31, nside=16, ra_deg=(10., 16.), dec_deg=(5., 11.),
                                     z_range=(1.9, 3.2), cosmo=cosmo)
    for qsos in objs.values():  # picca_xcf_angl.py: observed Lyman-alpha wavelength of the object
        for q in qsos:
            q.log_lambda = np.log10((1. + q.z_qso) * synth.LYA)
    return objs, z_min


def quasars2(cosmo):
    """second object catalogue of the crossed co cases (other seed, other ids)"""
    return synth.make_quasars(300, seed=37, nside=16, ra_deg=(10., 16.), dec_deg=(5., 11.),
                              z_range=(1.9, 3.2), cosmo=cosmo, id_offset=2 * 10**7)


def ang_max_for(cosmo, cfg, z_min, z_min2=None):
    if cfg.get("ang_correlation"):  # picca_cf_angl.py:214,222: ang_max IS r_trans_max
        return cfg["r_trans_max"]
    return synth.compute_ang_max(cosmo, cfg["r_trans_max"], z_min, z_min2)
