"""Snapshot of the module globals of ``picca.cf`` / ``picca.xcf`` taken at call time
(reference py/picca/cf.py:28-79, py/picca/xcf.py:27-68)."""
from ._lib import Params


def params_from_module(mod, cross=False):
    def g(name, default=None):
        return getattr(mod, name, default)

    for name in ("num_bins_r_par", "num_bins_r_trans", "r_par_min", "r_par_max", "r_trans_max"):
        if g(name) is None:
            raise RuntimeError("picca_b200: module global `%s` is not set" % name)
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
    p.has_zerr_cut = int(g("zerr_cut_deg") is not None)
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
    p.ang_max = float(g("ang_max") if g("ang_max") is not None else 0.0)
    return p
