"""GPU: the covariance step (picca_b200.export -> pb2_cov_subsample / pb2_cov_smooth through the
C ABI) against the oracle restatement of utils.compute_cov / utils.smooth_cov and against the live
reference's golden outputs (tests/golden/golden_export.npz).

Tolerance: the kernels re-associate the fp64 sums (tiled contraction, atomics), so entries agree
within 1e-9 of sqrt(var_i var_j) -- the north_star tolerance for fp64 sums on the scale of the
summed terms; the weighted means and weight sums are bit-equal (same association as NumPy)."""
import os

import numpy as np
import pytest

from tests.golden import cases_export

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_export.npz")
RTOL = 1e-9


def cov_close(got, want, rtol=RTOL):
    sd = np.sqrt(np.abs(np.diagonal(want)))
    scale = np.maximum(sd[:, None] * sd[None, :], np.abs(want))
    assert got.shape == want.shape
    err = np.abs(got - want)
    assert np.all(err <= rtol * scale + 1e-300), \
        "max scaled error %.3e" % np.max(err / np.maximum(scale, 1e-300))


@pytest.mark.parametrize("name", sorted(cases_export.CASES))
def test_cov_and_smoothing_match_reference_golden(name):
    from picca_b200 import export
    export.userprint = lambda *a, **k: None
    gold = np.load(GOLD)
    cfg = cases_export.CASES[name]
    xi, we, rp, rt = cases_export.inputs(cfg)
    cov = export.compute_cov(xi, we)
    cov_close(cov, gold["%s_cov" % name])
    assert np.array_equal(cov, cov.T)
    kw = dict(delta_r_trans=cfg["delta_r_trans"], delta_r_par=cfg["delta_r_par"],
              per_r_par=cfg.get("per_r_par", False))
    smooth = export.smooth_cov(xi, we, rp, rt, covariance=gold["%s_cov" % name], **kw)
    cov_close(smooth, gold["%s_smooth" % name])
    # covariance=None computes it first (utils.py:182-183)
    smooth2 = export.smooth_cov(xi, we, rp, rt, **kw)
    cov_close(smooth2, gold["%s_smooth" % name], rtol=1e-8)


@pytest.mark.parametrize("name", sorted(cases_export.CASES))
def test_bootstrap_covariance_matches_reference_golden(name):
    """utils.compute_cov_boot (utils.py:131-150): same realisations (the reference's generator and
    call sequence), means and np.cov on the device."""
    from picca_b200 import export
    gold = np.load(GOLD)
    xi, we, _, _ = cases_export.inputs(cases_export.CASES[name])
    got = export.compute_cov_boot(xi, we, nboots=cases_export.NBOOTS, seed=cases_export.BOOT_SEED)
    want = gold["%s_boot" % name]
    assert np.array_equal(np.isnan(got), np.isnan(want))  # empty bin: 0/0 in every realisation
    ok = ~np.isnan(want)
    cov_close(np.where(ok, got, 0.), np.where(ok, want, 0.))
    assert np.array_equal(got[ok], got.T[ok])


def test_bootstrap_covariance_production_shape():
    """1392 sub-samples x 2500 bins, 2000 realisations, against the oracle on a column subset
    (the oracle's per-realisation gather is what makes the reference slow)."""
    from oracle import export as oexp
    from picca_b200 import export
    cfg = dict(n_s=1392, np_=50, nt=50, delta_r_par=4., delta_r_trans=4., seed=12)
    xi, we, _, _ = cases_export.inputs(cfg)
    got = export.compute_cov_boot(xi, we, nboots=2000, seed=5)
    cols = np.arange(0, 2500, 41)
    want = oexp.compute_cov_boot(xi[:, cols], we[:, cols], nboots=2000, seed=5)
    cov_close(got[np.ix_(cols, cols)], want)


def test_reference_fixture_exported_cf():
    """cf.fits.gz -> exported_cf.fits.gz CO column (the reference's own golden, rtol 1e-5)."""
    from picca_b200 import export
    export.userprint = lambda *a, **k: None
    gold = np.load(GOLD)
    n_p, n_t, rp_min, rp_max, rt_max = gold["fixture_bins"]
    cov = export.compute_cov(gold["fixture_da"], gold["fixture_we"])
    smooth = export.smooth_cov(None, None, gold["fixture_rp"], gold["fixture_rt"],
                               delta_r_trans=(rt_max - 0.) / n_t,
                               delta_r_par=(rp_max - rp_min) / n_p, covariance=cov)
    np.testing.assert_allclose(smooth, gold["fixture_co"], rtol=1e-5,
                               atol=1e-8 * np.abs(gold["fixture_co"]).max())


def test_means_and_weight_sums_are_bit_equal():
    from picca_b200 import export
    from picca_b200.engine import get_engine
    eng = get_engine()
    xi, we, _, _ = cases_export.inputs(cases_export.CASES["ragged"])
    _, mean_xi, sum_w = export.compute_cov_device(eng, export._dev(eng, xi), export._dev(eng, we))
    want_w = we.sum(axis=0)
    want_m = (xi * we).sum(axis=0)
    want_m[want_w > 0] /= want_w[want_w > 0]
    assert np.array_equal(sum_w.cpu().numpy(), want_w)
    assert np.array_equal(mean_xi.cpu().numpy(), want_m)


@pytest.mark.parametrize("n_s,np_,nt", [(1, 3, 3), (2, 1, 1), (16, 8, 8), (17, 13, 5), (50, 65, 1)])
def test_ragged_shapes_against_oracle(n_s, np_, nt):
    """tile-edge shapes: one sub-sample, one bin, exact multiples of the 64-bin tile / 16-sample
    chunk and one past them"""
    from oracle import export as oexp
    from picca_b200 import export
    export.userprint = lambda *a, **k: None
    cfg = dict(n_s=n_s, np_=np_, nt=nt, delta_r_par=4., delta_r_trans=4., seed=100 + n_s)
    xi, we, rp, rt = cases_export.inputs(cfg)
    want = oexp.compute_cov(xi, we)
    got = export.compute_cov(xi, we)
    cov_close(got, want)
    if n_s > 1 and np_ * nt > 1:
        cov_close(export.smooth_cov(xi, we, rp, rt, covariance=want),
                  oexp.smooth_cov(xi, we, rp, rt, covariance=want))


def test_production_size_against_oracle_and_properties():
    """config-2 shape: 1392 HEALPix sub-samples x 2500 bins, and the xcf default of 5000 bins on a
    smaller sample.  Checked against the oracle, plus size-independent properties: symmetry,
    non-negative variances, invariance under a permutation of the sub-samples, and the Cauchy-
    Schwarz bound of a covariance."""
    from oracle import export as oexp
    from picca_b200 import export
    export.userprint = lambda *a, **k: None
    for n_s, np_, nt in ((1392, 50, 50), (300, 100, 50)):
        cfg = dict(n_s=n_s, np_=np_, nt=nt, delta_r_par=4., delta_r_trans=4., seed=11)
        xi, we, rp, rt = cases_export.inputs(cfg)
        got = export.compute_cov(xi, we)
        cov_close(got, oexp.compute_cov(xi, we))
        assert np.array_equal(got, got.T)
        var = np.diagonal(got)
        assert np.all(var > 0)
        assert np.all(np.abs(got) <= np.sqrt(var[:, None] * var[None, :]) * (1 + 1e-12))
        perm = np.random.default_rng(0).permutation(n_s)
        cov_close(export.compute_cov(xi[perm], we[perm]), got, rtol=1e-10)
        smooth = export.smooth_cov(xi, we, rp, rt, covariance=got)
        assert np.array_equal(smooth, smooth.T)
        np.testing.assert_allclose(np.diagonal(smooth), var, rtol=1e-15)
        if np_ * nt <= 2500:
            cov_close(smooth, oexp.smooth_cov(xi, we, rp, rt, covariance=got))
