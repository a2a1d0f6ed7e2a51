"""ctypes binding of ``libpicca_b200.so`` -- the C ABI declared in ``include/picca_b200.h``.

There is NO fallback: if the CUDA library is missing or does not load, importing this module's
``lib()`` raises, and so does every product entry point built on it.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PICCA_B200_LIB", os.path.join(_HERE, "libpicca_b200.so"))
_LIB = None

c_i32p = ctypes.c_void_p
c_void_p = ctypes.c_void_p


class Params(ctypes.Structure):
    """``pb2_params``: the picca.cf / picca.xcf module globals (cf.py:28-79, xcf.py:27-68)."""
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
        ("ang_max", ctypes.c_double),
    ]


class Catalog(ctypes.Structure):
    """``pb2_catalog``: device pointers of a packed catalogue."""
    _fields_ = [
        ("n_los", ctypes.c_int64),
        ("n_pix", ctypes.c_int64),
        ("offset", c_void_p),
        ("r_comov", c_void_p),
        ("dist_m", c_void_p),
        ("z", c_void_p),
        ("weights", c_void_p),
        ("delta_w", c_void_p),
        ("z_w", c_void_p),
        ("log_lambda", c_void_p),
        ("dg_offset", c_void_p),
        ("dg_count", c_void_p),
        ("dg_rec", c_void_p),
        ("il_offset", c_void_p),
        ("il_total", ctypes.c_int64),
        ("il_rec", c_void_p),
        ("dg_lanes", ctypes.c_int32),
        ("dg_max_pix", ctypes.c_int32),
        ("dg_ok", ctypes.c_int32),
        ("dg_reserved", ctypes.c_int32),
        ("dg_reach", ctypes.c_double),
        ("px_rec", c_void_p),
        ("x_cart", c_void_p),
        ("y_cart", c_void_p),
        ("z_cart", c_void_p),
        ("ra", c_void_p),
        ("dec", c_void_p),
        ("cos_dec", c_void_p),
        ("z_qso", c_void_p),
        ("thingid", c_void_p),
        ("plate", c_void_p),
        ("fiberid", c_void_p),
        ("order", c_void_p),
        ("row", c_void_p),
        ("n_hp", ctypes.c_int32),
        ("sorted", ctypes.c_int32),
        ("hp_first", c_void_p),
        ("cap_x", c_void_p),
        ("cap_y", c_void_p),
        ("cap_z", c_void_p),
        ("cap_rad", c_void_p),
        ("max_pix", ctypes.c_int32),
        ("reserved", ctypes.c_int32),
    ]


class Pairs(ctypes.Structure):
    """``pb2_pairs``: device pointers of a CSR forest-pair list."""
    _fields_ = [
        ("n_f1", ctypes.c_int64),
        ("n_pairs", ctypes.c_int64),
        ("f1_index", c_void_p),
        ("nb_offset", c_void_p),
        ("nb_f1", c_void_p),
        ("nb_f2", c_void_p),
        ("nb_ang", c_void_p),
        ("nb_cos", c_void_p),
        ("nb_sin", c_void_p),
        ("nb_keep", c_void_p),
    ]


# every symbol include/picca_b200.h declares (checked by tests/test_abi.py)
EXPORTS = [
    "pb2_abi_version", "pb2_last_error", "pb2_sizeof_params", "pb2_sizeof_catalog",
    "pb2_sizeof_pairs", "pb2_diag_lanes", "pb2_neigh_count", "pb2_neigh_fill", "pb2_xi_auto", "pb2_xi_cross",
    "pb2_pack_diag", "pb2_catalog_stats", "pb2_derive_products", "pb2_build_prefix", "pb2_xi_normalise", "pb2_dmat_scratch_bytes", "pb2_dmat_auto", "pb2_dmat_cross", "pb2_dmat_stats",
    "pb2_metal_dmat_auto", "pb2_metal_dmat_cross", "pb2_wick_scratch_bytes", "pb2_wick_auto",
    "pb2_wick_cross", "pb2_co_pairs", "pb2_cov_scratch_bytes", "pb2_cov_subsample", "pb2_cov_smooth",
    "pb2_cov_boot_scratch_bytes", "pb2_cov_boot",
    "pb2_fits_scan", "pb2_fits_cards", "pb2_delta_unpack", "pb2_delta_prepare",
    "pb2_delta_image_count", "pb2_delta_image_unpack",
    "pb2_fits_hierarch", "pb2_delta_wave", "pb2_delta_rebin",
    "pb2_fp64_peak", "pb2_launch_count", "pb2_set_timing", "pb2_last_kernel_ms",
]

ABI_VERSION = 23


def lib():
    """Load the CUDA library (once).  Raises RuntimeError when it is absent -- never falls back."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "picca_b200: %s is missing. Build it with `python -m picca_b200.csrc.build` "
            "(or __graft_entry__.build()); there is no CPU fallback." % LIB_PATH)
    handle = ctypes.CDLL(LIB_PATH)
    handle.pb2_last_error.restype = ctypes.c_char_p
    handle.pb2_launch_count.restype = ctypes.c_int64
    handle.pb2_last_kernel_ms.restype = ctypes.c_double
    handle.pb2_dmat_scratch_bytes.restype = ctypes.c_int64
    handle.pb2_cov_scratch_bytes.restype = ctypes.c_int64
    handle.pb2_fits_scan.restype = ctypes.c_int64
    handle.pb2_cov_boot_scratch_bytes.restype = ctypes.c_int64
    handle.pb2_wick_scratch_bytes.restype = ctypes.c_int64
    if handle.pb2_abi_version() != ABI_VERSION:
        raise RuntimeError("picca_b200: ABI mismatch, rebuild libpicca_b200.so")
    assert handle.pb2_sizeof_params() == ctypes.sizeof(Params)
    assert handle.pb2_sizeof_catalog() == ctypes.sizeof(Catalog)
    assert handle.pb2_sizeof_pairs() == ctypes.sizeof(Pairs)
    _LIB = handle
    return _LIB


def check(status, what):
    """Turn a non-zero status into a Python exception (CUDA errors never pass silently)."""
    if status != 0:
        msg = lib().pb2_last_error().decode("utf-8", "replace")
        raise RuntimeError("picca_b200: %s failed (status %d): %s" % (what, status, msg))
