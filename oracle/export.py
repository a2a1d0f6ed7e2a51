"""Oracle restatement of the covariance step of ``picca_export.py``: ``utils.compute_cov``
(reference py/picca/utils.py:100-128) and ``utils.smooth_cov`` (py/picca/utils.py:153-249).
TEST INFRASTRUCTURE ONLY -- the referee for picca_b200.export, never the product.

Pinned (tests/test_oracle_vs_reference.py, tests/test_oracle_golden.py) against the live
reference functions on seeded inputs, against tests/golden/golden_export.npz (their outputs,
committed) and against the reference's own golden ``exported_cf.fits.gz`` (CO column from
``cf.fits.gz``, rtol 1e-5 as the reference's test compares).
"""
import numpy as np


def compute_cov(xi, weights):
    """utils.py:113-128, statement by statement (NumPy: axis-0 sums add the rows in order)."""
    xi = np.asarray(xi, dtype=np.float64)
    weights = np.asarray(weights, dtype=np.float64)
    mean_xi = (xi * weights).sum(axis=0)                      # :113
    sum_weights = weights.sum(axis=0)                         # :114
    ok = sum_weights > 0.                                     # :115
    mean_xi[ok] /= sum_weights[ok]                            # :116
    m = weights * (xi - mean_xi)                              # :118
    covariance = m.T.dot(m)                                   # :122
    denom = sum_weights * sum_weights[:, None]                # :123
    ok = denom > 0.                                           # :124
    covariance[ok] /= denom[ok]                               # :125
    return covariance


def compute_cov_boot(xi, weights, nboots=10000, seed=121567):
    """utils.py:143-150"""
    xi = np.asarray(xi, dtype=np.float64)
    weights = np.asarray(weights, dtype=np.float64)
    nhpx, ndata = xi.shape
    boot_xis = np.empty((nboots, ndata))
    rnst = np.random.default_rng(seed)                        # :145
    for i in range(nboots):
        idx = rnst.choice(nhpx, size=nhpx)                    # :148
        wei = weights[idx]
        boot_xis[i] = np.sum(wei * xi[idx], axis=0) / wei.sum(0)   # :150
    return np.cov(boot_xis, rowvar=False)


def smooth_cov(xi, weights, r_par, r_trans, delta_r_trans=4.0, delta_r_par=4.0, covariance=None,
               per_r_par=False):
    """utils.py:182-248.  The reference's two Python loops over (index, index2 > index) are
    restated with the pairs enumerated in the same row-major order and ``np.add.at`` (sequential,
    unbuffered), so every dictionary sum is formed in the reference's order."""
    if covariance is None:
        covariance = compute_cov(xi, weights)                 # :182-183
    covariance = np.asarray(covariance, dtype=np.float64)
    r_par = np.asarray(r_par, dtype=np.float64)
    r_trans = np.asarray(r_trans, dtype=np.float64)
    num_bins = covariance.shape[1]
    var = np.diagonal(covariance)
    if np.any(var == 0.):                                     # :187-190
        return covariance
    correlation = covariance / np.sqrt(var * var[:, None])    # :192
    i, j = np.triu_indices(num_bins, 1)                       # index, index2 in loop order
    # round() of a Python float rounds half to even, like np.rint           :207-210
    k_dp = np.rint(np.abs(r_par[j] - r_par[i]) / delta_r_par).astype(np.int64)
    k_dt = np.rint(np.abs(r_trans[i] - r_trans[j]) / delta_r_trans).astype(np.int64)
    key = k_dp * (k_dt.max() + 1) + k_dt
    if per_r_par:
        k_rp = np.trunc(r_par[i] / delta_r_par).astype(np.int64)  # int(): towards zero   :204
        k_rp -= k_rp.min()
        key = key + k_rp * ((k_dp.max() + 1) * (k_dt.max() + 1))
    total = np.zeros(key.max() + 1 if key.size else 1)
    count = np.zeros(key.max() + 1 if key.size else 1, dtype=np.int64)
    np.add.at(total, key, correlation[i, j])                  # :217-229
    np.add.at(count, key, 1)
    correlation_smooth = np.zeros([num_bins, num_bins])
    correlation_smooth[np.arange(num_bins), np.arange(num_bins)] = 1.   # :232
    correlation_smooth[i, j] = total[key] / count[key]        # :240-245
    correlation_smooth[j, i] = correlation_smooth[i, j]       # :246-247
    return correlation_smooth * np.sqrt(var * var[:, None])   # :250-251
