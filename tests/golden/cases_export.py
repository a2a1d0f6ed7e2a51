"""Seeded inputs of the covariance golden vectors (shared by the generator and the tests)."""
import numpy as np

NBOOTS, BOOT_SEED = 70, 121567  # bootstrap realisations of the golden vectors (utils.py:131)

CASES = {
    # name: sub-samples, np, nt, bin widths, options
    "small": dict(n_s=23, np_=6, nt=5, delta_r_par=4., delta_r_trans=4., seed=3),
    "ragged": dict(n_s=17, np_=9, nt=7, delta_r_par=4., delta_r_trans=4., seed=4, empty_rows=3),
    "per_r_par": dict(n_s=40, np_=8, nt=8, delta_r_par=4., delta_r_trans=4., seed=5,
                      per_r_par=True),
    "xcf_like": dict(n_s=31, np_=10, nt=5, delta_r_par=4., delta_r_trans=4., seed=6,
                     rp_min=-20., per_r_par=True),
    "empty_bin": dict(n_s=12, np_=4, nt=4, delta_r_par=4., delta_r_trans=4., seed=7,
                      empty_bin=5),
}


def inputs(cfg):
    """xi [n_s, nb], weights [n_s, nb], r_par [nb], r_trans [nb] like a picca_cf.py output:
    bin centres jittered like weighted means, weights varying by a factor of a few between
    sub-samples, a common signal plus noise."""
    rng = np.random.default_rng(cfg["seed"])
    n_s, np_, nt = cfg["n_s"], cfg["np_"], cfg["nt"]
    nb = np_ * nt
    rp_min = cfg.get("rp_min", 0.)
    bp, bt = np.divmod(np.arange(nb), nt)
    r_par = rp_min + (bp + 0.5 + 0.2 * rng.uniform(-1, 1, nb)) * cfg["delta_r_par"]
    r_trans = (bt + 0.5 + 0.2 * rng.uniform(-1, 1, nb)) * cfg["delta_r_trans"]
    signal = 1e-3 * np.cos(0.1 * r_par) / (1. + 0.05 * r_trans)
    weights = rng.uniform(0.5, 3., (n_s, nb)) * rng.uniform(10., 1000., (n_s, 1))
    common = rng.normal(0., 1., (n_s, 1)) * 2e-4   # correlated part
    xi = signal + common + rng.normal(0., 1., (n_s, nb)) / np.sqrt(weights)
    for k in range(cfg.get("empty_rows", 0)):   # sub-samples without pairs in some bins
        weights[k, rng.integers(0, nb, nb // 3)] = 0.
    if "empty_bin" in cfg:                       # a bin nobody filled: zero variance
        weights[:, cfg["empty_bin"]] = 0.
        xi[:, cfg["empty_bin"]] = 0.
    return xi, weights, r_par, r_trans
