"""B200 path of the covariance step of ``picca_export.py``: the sub-sample covariance of the
per-HEALPix correlation blocks and its smoothing.  Same names, arguments and return values as the
reference's ``picca.utils.compute_cov`` (py/picca/utils.py:100-128), ``compute_cov_boot``
(py/picca/utils.py:131-150) and ``picca.utils.smooth_cov`` (py/picca/utils.py:153-249);
``picca_b200.overlay`` makes ``picca.utils`` resolve them here so the unmodified
``picca_export.py`` (:262-300) runs on top.

NumPy host arrays in, NumPy host arrays out (what the script holds); the arithmetic runs in
``pb2_cov_subsample`` / ``pb2_cov_boot`` / ``pb2_cov_smooth`` (csrc/pb2_cov.cu) through the C ABI.  No CPU fallback.
"""
import ctypes
import sys

import numpy as np

from . import _lib
from .engine import get_engine


def userprint(*args, **kwds):
    """reference py/picca/utils.py:31-40"""
    print(*args, **kwds)
    sys.stdout.flush()


def _dev(eng, a):
    return eng.torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(eng.device)


def compute_cov_device(eng, d_xi, d_weights):
    """Device tensors [n_samples, nb] -> (cov [nb, nb], mean_xi [nb], sum_weights [nb]) on the
    device."""
    torch = eng.torch
    n_s, nb = int(d_xi.shape[0]), int(d_xi.shape[1])
    cov = torch.empty((nb, nb), dtype=torch.float64, device=eng.device)
    mean_xi = torch.empty(nb, dtype=torch.float64, device=eng.device)
    sum_w = torch.empty(nb, dtype=torch.float64, device=eng.device)
    nbytes = int(eng.lib.pb2_cov_scratch_bytes(ctypes.c_int64(n_s), ctypes.c_int32(nb)))
    scratch = torch.empty(nbytes, dtype=torch.uint8, device=eng.device)
    _lib.check(eng.lib.pb2_cov_subsample(
        ctypes.c_int64(n_s), ctypes.c_int32(nb), ctypes.c_void_p(d_xi.data_ptr()),
        ctypes.c_void_p(d_weights.data_ptr()), ctypes.c_void_p(cov.data_ptr()),
        ctypes.c_void_p(mean_xi.data_ptr()), ctypes.c_void_p(sum_w.data_ptr()),
        ctypes.c_void_p(scratch.data_ptr()), ctypes.c_int64(nbytes), eng.stream_ptr()),
        "pb2_cov_subsample")
    return cov, mean_xi, sum_w


def compute_cov(xi, weights):
    """Computes the covariance matrix using the subsampling technique (utils.py:100-128).

    Args:
        xi: array [n_healpix, nb] -- correlation function measurement in each healpix
        weights: array [n_healpix, nb] -- weights on the correlation function measurement
    Returns:
        The covariance matrix [nb, nb]
    """
    xi = np.asarray(xi, dtype=np.float64)
    weights = np.asarray(weights, dtype=np.float64)
    if xi.ndim != 2 or xi.shape != weights.shape:
        raise ValueError("compute_cov: xi and weights must be 2-D arrays of the same shape")
    eng = get_engine()
    userprint("Computing cov...")
    cov, _, _ = compute_cov_device(eng, _dev(eng, xi), _dev(eng, weights))
    return cov.cpu().numpy()


def compute_cov_boot(xi, weights, nboots=10000, seed=121567):
    """Computes the covariance matrix using the bootstrap technique (utils.py:131-150).

    The resampling indices come from the reference's generator and call sequence
    (``np.random.default_rng(seed)``, one ``choice(nhpx, size=nhpx)`` per realisation), so the
    realisations are the reference's; the weighted means and ``np.cov`` run on the device.
    """
    xi = np.asarray(xi, dtype=np.float64)
    weights = np.asarray(weights, dtype=np.float64)
    if xi.ndim != 2 or xi.shape != weights.shape:
        raise ValueError("compute_cov_boot: xi and weights must be 2-D arrays of the same shape")
    nhpx, ndata = xi.shape
    eng = get_engine()
    torch = eng.torch
    rnst = np.random.default_rng(seed)
    idx = np.empty((nboots, nhpx), dtype=np.int32)
    for i in range(nboots):
        idx[i] = rnst.choice(nhpx, size=nhpx)
    d_idx = torch.from_numpy(idx).to(eng.device)
    d_xi, d_we = _dev(eng, xi), _dev(eng, weights)
    cov = torch.empty((ndata, ndata), dtype=torch.float64, device=eng.device)
    nbytes = int(eng.lib.pb2_cov_boot_scratch_bytes(ctypes.c_int64(nhpx), ctypes.c_int32(ndata),
                                                    ctypes.c_int32(nboots)))
    scratch = torch.empty(nbytes, dtype=torch.uint8, device=eng.device)
    _lib.check(eng.lib.pb2_cov_boot(
        ctypes.c_int64(nhpx), ctypes.c_int32(ndata), ctypes.c_int32(nboots),
        ctypes.c_void_p(d_xi.data_ptr()), ctypes.c_void_p(d_we.data_ptr()),
        ctypes.c_void_p(d_idx.data_ptr()), ctypes.c_void_p(cov.data_ptr()),
        ctypes.c_void_p(scratch.data_ptr()), ctypes.c_int64(nbytes), eng.stream_ptr()),
        "pb2_cov_boot")
    return cov.cpu().numpy()


def _smooth_extents(r_par, r_trans, delta_r_par, delta_r_trans, per_r_par):
    """Extents of the reference's dictionary keys (utils.py:207-211): the largest rounded
    differences, and the range of int(r_par/delta) when smoothing per r_par."""
    n_dp = int(round(abs(float(np.max(r_par)) - float(np.min(r_par))) / delta_r_par)) + 2
    n_dt = int(round(abs(float(np.max(r_trans)) - float(np.min(r_trans))) / delta_r_trans)) + 2
    rp_lo, n_rp = 0, 1
    if per_r_par:
        rp_lo = int(float(np.min(r_par)) / delta_r_par)
        n_rp = int(float(np.max(r_par)) / delta_r_par) - rp_lo + 1
    return n_dp, n_dt, rp_lo, n_rp


def smooth_cov(xi, weights, r_par, r_trans, delta_r_trans=4.0, delta_r_par=4.0, covariance=None,
               per_r_par=False):
    """Smoothes the covariance matrix (utils.py:153-249): the correlation coefficient of two bins
    is replaced by its mean over all bin pairs with the same rounded separation differences.

    If the data has empty bins (a zero variance) prints the reference's warnings and returns the
    unsmoothed covariance (utils.py:187-190).
    """
    eng = get_engine()
    torch = eng.torch
    if covariance is None:
        d_cov, _, _ = compute_cov_device(eng, _dev(eng, xi), _dev(eng, weights))
    else:
        d_cov = _dev(eng, covariance)
    num_bins = int(d_cov.shape[1])
    var = torch.diagonal(d_cov).cpu().numpy()
    if np.any(var == 0.):
        userprint('WARNING: data has some empty bins, impossible to smooth')
        userprint('WARNING: returning the unsmoothed covariance')
        return d_cov.cpu().numpy() if covariance is None else covariance
    r_par = np.ascontiguousarray(r_par, dtype=np.float64)
    r_trans = np.ascontiguousarray(r_trans, dtype=np.float64)
    if r_par.size != num_bins or r_trans.size != num_bins:
        raise ValueError("smooth_cov: r_par / r_trans do not match the covariance")
    n_dp, n_dt, rp_lo, n_rp = _smooth_extents(r_par, r_trans, delta_r_par, delta_r_trans,
                                              per_r_par)
    keys = n_rp * n_dp * n_dt
    tab_sum = torch.empty(keys, dtype=torch.float64, device=eng.device)
    tab_cnt = torch.empty(keys, dtype=torch.int64, device=eng.device)
    bad = torch.zeros(1, dtype=torch.int32, device=eng.device)
    out = torch.empty_like(d_cov)
    d_rp, d_rt = _dev(eng, r_par), _dev(eng, r_trans)
    _lib.check(eng.lib.pb2_cov_smooth(
        ctypes.c_int32(num_bins), ctypes.c_void_p(d_cov.data_ptr()),
        ctypes.c_void_p(d_rp.data_ptr()), ctypes.c_void_p(d_rt.data_ptr()),
        ctypes.c_double(delta_r_par), ctypes.c_double(delta_r_trans),
        ctypes.c_int32(int(bool(per_r_par))), ctypes.c_int32(n_dp), ctypes.c_int32(n_dt),
        ctypes.c_int32(rp_lo), ctypes.c_int32(n_rp), ctypes.c_void_p(tab_sum.data_ptr()),
        ctypes.c_void_p(tab_cnt.data_ptr()), ctypes.c_void_p(bad.data_ptr()),
        ctypes.c_void_p(out.data_ptr()), eng.stream_ptr()), "pb2_cov_smooth")
    if int(bad.item()):
        raise RuntimeError("picca_b200: smooth_cov key outside the precomputed extents "
                           "(non-finite r_par / r_trans?)")
    userprint("\n")
    return out.cpu().numpy()
