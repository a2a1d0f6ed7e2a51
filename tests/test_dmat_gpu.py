"""GPU parity of the distortion matrix: picca_b200.cf.compute_dmat / xcf.compute_dmat (CUDA)
against the live reference's golden vectors.  NPALL / NPUSED (pair counts under the --rej draw)
must be exact; the matrix and the effective quantities within 1e-9 relative (north_star), with an
absolute floor of 1e-12 x the largest entry for elements that are sums of cancelling terms."""
import os

import numpy as np
import pytest

from tests import helpers
from tests.golden import cases

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["run", "dense"], autouse=True)
def dmat_kernel(request, monkeypatch):
    """Every test runs on both auto-correlation kernels: the product one (run lengths + prefix
    sums, register tile; pb2_dmat_run.cu) and the dense-scratch one kept for r-mu binning
    (pb2_dmat.cu), selected per call through PB2_DMAT_KERNEL."""
    monkeypatch.setenv("PB2_DMAT_KERNEL", request.param)
    return request.param


GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NAMES = ("weights_dmat", "dmat", "r_par_eff", "r_trans_eff", "z_eff", "weight_eff")


def check8(res, gold, prefix):
    assert [int(res[6]), int(res[7])] == list(gold[prefix + "counts"])
    for k, n in enumerate(NAMES):
        want = gold[prefix + n]
        got = np.asarray(res[k])
        assert got.shape == want.shape
        scale = np.abs(want).max()
        err = np.abs(got - want)
        tol = 1e-9 * np.abs(want) + 1e-12 * scale
        assert np.all(err <= tol), "%s%s: max err %.3e (scale %.3e), worst rel %.3e" % (
            prefix, n, err.max(), scale, (err / np.maximum(np.abs(want), 1e-300))[err > tol].max())
        assert np.array_equal(got != 0, want != 0) or n != "weights_dmat"


@pytest.mark.parametrize("name", sorted(cases.DMAT_CASES))
def test_dmat_matches_reference_golden(name):
    from picca_b200 import cf
    gold = np.load(os.path.join(GOLD, "golden_dmat.npz"))
    cfg = dict(cases.DMAT_CASES[name])
    second = cfg.pop("second", False)
    data, num, z_min, cosmo = cases.dmat_forests()
    over, z_min2 = dict(cfg), None
    if second:
        data2, num2, z_min2, _ = cases.dmat_forests(second=True)
        over["data2"], over["num_data2"] = data2, num2
    helpers.configure(cf, data, num, cases.ang_max_for(cosmo, cfg, z_min, z_min2), **over)
    hps = sorted(data)
    cf.fill_neighs(hps)
    np.random.seed(hps[0])  # picca_dmat.py:36
    res = cf.compute_dmat(hps)
    check8(res, gold, "dmat_%s_" % name)


@pytest.mark.parametrize("name", sorted(cases.XDMAT_CASES))
def test_xdmat_matches_reference_golden(name):
    from picca_b200 import xcf
    gold = np.load(os.path.join(GOLD, "golden_xdmat.npz"))
    cfg = cases.XDMAT_CASES[name]
    data, num, z_min, cosmo = cases.dmat_forests()
    objs, z_min2 = cases.quasars(cosmo)
    helpers.configure(xcf, data, num, cases.ang_max_for(cosmo, cfg, z_min, z_min2), objs=objs,
                      **cfg)
    hps = sorted(data)
    xcf.fill_neighs(hps)
    np.random.seed(hps[0])  # picca_xdmat.py:37
    res = xcf.compute_dmat(hps)
    check8(res, gold, "xdmat_%s_" % name)


def test_dmat_chunks_sum_like_the_script():
    """picca_dmat.py sums the 8-tuples of its workers (:494-501): two chunks == one chunk when the
    same pairs are kept (reject = 0 keeps every pair, no RNG dependence)."""
    from picca_b200 import cf
    cfg = dict(cases.DMAT_CASES["default"], reject=0.)
    data, num, z_min, cosmo = cases.dmat_forests()
    data = {hp: v[:6] for hp, v in data.items()}
    helpers.configure(cf, data, num, cases.ang_max_for(cosmo, cfg, z_min), **cfg)
    hps = sorted(data)
    cf.fill_neighs(hps)
    whole = cf.compute_dmat(hps)
    half = len(hps) // 2
    cf.fill_neighs(hps[:half])
    a = cf.compute_dmat(hps[:half])
    cf.fill_neighs(hps[half:])
    b = cf.compute_dmat(hps[half:])
    assert whole[6] == a[6] + b[6] and whole[7] == a[7] + b[7] == whole[6]
    for k in range(6):
        s = a[k] + b[k]
        np.testing.assert_allclose(s, whole[k], rtol=1e-9, atol=1e-12 * np.abs(whole[k]).max())


def _dmat_vs_oracle(data, num, ang_max, **cfg):
    from oracle import cf as ocf
    from picca_b200 import cf
    helpers.configure(ocf, data, num, ang_max, **cfg)
    helpers.configure(cf, data, num, ang_max, **cfg)
    hps = sorted(data)
    ocf.fill_neighs(hps)
    np.random.seed(hps[0])
    want = ocf.compute_dmat(hps)
    cf.fill_neighs(hps)
    np.random.seed(hps[0])
    got = cf.compute_dmat(hps)
    assert (int(got[6]), int(got[7])) == (int(want[6]), int(want[7]))
    for k, n in enumerate(NAMES):
        w, g = np.asarray(want[k]), np.asarray(got[k])
        err = np.abs(g - w)
        tol = 1e-9 * np.abs(w) + 1e-12 * np.abs(w).max()
        assert np.all(err <= tol), "%s: max err %.3e (scale %.3e)" % (n, err.max(), np.abs(w).max())
    return int(want[7])


def test_dmat_production_binning_matches_oracle():
    """50 x 50 bins, r < 200: several hundred touched bins per forest pair (compact-bin chunks,
    r_trans-major tiles and the tile skipping of the contraction)."""
    from picca_b200 import synth
    data, num, z_min, _, cosmo = helpers.small_sample(n=60, seed=91, max_pix=260, side_deg=3.)
    ang_max = synth.compute_ang_max(cosmo, 200., z_min)
    used = _dmat_vs_oracle(data, num, ang_max, num_bins_r_par=50, num_bins_r_trans=50,
                           num_model_bins_r_par=50, num_model_bins_r_trans=50, r_par_max=200.,
                           r_trans_max=200., reject=0.9)
    assert used > 20


def test_dmat_long_forests_without_tile_skipping():
    """Forests of ~2200 pixels: more K chunks of rows than the kernel keeps column ranges for, so
    the contraction walks every chunk."""
    from picca_b200 import synth
    data, num, z_min, _, cosmo = synth.make_forests(
        5, seed=6, nside=16, ra_deg=(10., 10.6), dec_deg=(5., 5.6), rest_range=(1000., 1250.),
        dlambda=0.25)
    lens = [len(d.weights) for hp in data for d in data[hp]]
    assert 2 * sorted(lens)[-1] + 2 * sorted(lens)[-2] > 256 * 32 > 4 * min(lens)
    ang_max = synth.compute_ang_max(cosmo, 60., z_min)
    used = _dmat_vs_oracle(data, num, ang_max, reject=0.)
    assert used >= 3


def test_dmat_same_half_plate_close_pairs_truncated_unique(dmat_kernel):
    """with --remove-same-half-plate-close-pairs the reference's pass 0 does not count
    the close pairs of a same-half-plate forest pair (cf.py:565-568) while pass 1 records the model
    bin of EVERY in-range pair in an array sized by that count (cf.py:702-703), so np.unique
    (cf.py:846-848) only sees the model bins of the first `count` in-range pairs.  The oracle
    restates exactly that ("record only while in bounds"); the product kernel reproduces it."""
    if dmat_kernel == "dense":
        pytest.skip("the dense-scratch kernel (r-mu binning only) does not emulate Q8")
    data, num, z_min, cosmo = cases.dmat_forests()
    cases.share_plates(data)
    cfg = dict(cases.DMAT_CASES["default"], reject=0.5, remove_same_half_plate_close_pairs=True)
    # there are same-half-plate neighbours, and they share a wavelength grid: close pairs exist
    plates = {}
    for v in data.values():
        for d in v:
            plates.setdefault((d.plate, d.fiberid <= 500), []).append(d)
    assert max(len(v) for v in plates.values()) > 5
    used = _dmat_vs_oracle(data, num, cases.ang_max_for(cosmo, cfg, z_min), **cfg)
    assert used > 100


def test_dmat_no_selected_pair_and_three_pixel_forests(dmat_kernel):
    """z-pair cuts that deselect everything (in-range pairs still feed the eta terms, nothing is
    added), and forests of three pixels (shorter than a warp step)."""
    from picca_b200 import synth
    data, num, z_min, _, cosmo = helpers.small_sample(n=60, seed=5, max_pix=3, side_deg=2.,
                                                      zero_weight_frac=0.)
    ang_max = synth.compute_ang_max(cosmo, 60., z_min)
    used = _dmat_vs_oracle(data, num, ang_max, reject=0.)
    assert used > 10
    data, num, z_min, cosmo = cases.dmat_forests()
    from oracle import cf as ocf
    from picca_b200 import cf
    cfg = dict(cases.DMAT_CASES["default"], reject=0.5, z_min_pairs=9., z_max_pairs=10.)
    for mod in (ocf, cf):
        helpers.configure(mod, data, num, cases.ang_max_for(cosmo, cfg, z_min), **cfg)
    hps = sorted(data)
    cf.fill_neighs(hps)
    np.random.seed(3)
    got = cf.compute_dmat(hps)
    assert got[7] > 100 and not got[1].any() and not got[0].any() and not got[5].any()
