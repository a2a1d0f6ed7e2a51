"""GPU parity of the Wick-expansion kernels (pb2_wick.cu) behind cf.compute_wick_terms (T1-T3,
reference cf.py:1326-1626) and xcf.compute_wick_terms (T1-T4, xcf.py:838-1351): against the
golden vectors of the LIVE reference (tests/golden/golden_wick.npz, golden_xwick.npz) and against
the oracle's C restatement on other seeded inputs.  num_pairs_wick / counts exact, fp64 sums
within 1e-9 of the matrix scale (entries are sums of signed products)."""
import os

import numpy as np
import pytest

from tests import helpers
from tests.golden import cases

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def close(got, want, tag):
    got, want = np.asarray(got), np.asarray(want)
    scale = np.abs(want).max()
    assert scale > 0, tag
    err = np.abs(got - want).max()
    assert err <= 1e-9 * scale, "%s: max abs err %.3e vs scale %.3e" % (tag, err, scale)
    assert np.array_equal(got != 0, want != 0), tag + ": touched cells differ"


def setup_wick(mod, cfg):
    cfg = dict(cfg)
    second = cfg.pop("second", False)
    data, num, z_min, cosmo = cases.dmat_forests()
    cases.set_fname(data, "D1")
    var1, xi1 = cases.wick_1d("D1")
    over = dict(cfg, get_variance_1d={"D1": var1}, xi_1d={"D1": xi1})
    z_min2 = None
    if second:
        data2, num2, z_min2, _ = cases.dmat_forests(second=True)
        over["data2"], over["num_data2"] = cases.set_fname(data2, "D2"), num2
        var2, xi2 = cases.wick_1d("D2")
        over["get_variance_1d"]["D2"], over["xi_1d"]["D2"] = var2, xi2
    helpers.configure(mod, data, num, cases.ang_max_for(cosmo, cfg, z_min, z_min2), **over)
    for k, v in over.items():
        setattr(mod, k, v)
    return data


@pytest.mark.parametrize("host_angles", [False, True])
@pytest.mark.parametrize("name", sorted(cases.WICK_CASES))
def test_wick_matches_reference_golden(name, host_angles, monkeypatch):
    from picca_b200 import _corr, cf
    monkeypatch.setattr(_corr, "HOST_ANGLES", host_angles)
    gold = np.load(os.path.join(GOLD, "golden_wick.npz"))
    data = setup_wick(cf, cases.WICK_CASES[name])
    hps = sorted(data)
    try:
        cf.fill_neighs(hps)
        np.random.seed(hps[0])  # picca_wick.py:35
        res = cf.compute_wick_terms(hps)
    finally:
        cf.data2 = None
    pre = "wick_%s_" % name
    assert (res[2], res[3]) == tuple(gold[pre + "counts"])
    assert np.array_equal(res[1], gold[pre + "num_pairs_wick"]) and res[1].sum() > 10000
    np.testing.assert_allclose(res[0], gold[pre + "weights_wick"], rtol=1e-9)
    for k in (1, 2, 3):
        close(res[3 + k], gold[pre + "t%d" % k], "%s t%d" % (name, k))
    assert not res[7].any() and not res[8].any() and not res[9].any()


@pytest.mark.parametrize("name", sorted(cases.XWICK_CASES))
def test_xwick_matches_reference_golden(name):
    from picca_b200 import xcf
    gold = np.load(os.path.join(GOLD, "golden_xwick.npz"))
    cfg = cases.XWICK_CASES[name]
    data, num, z_min, cosmo = cases.xwick_forests()
    cases.set_fname(data, "D1")
    objs, z_min2 = cases.quasars(cosmo)
    var1, xi1 = cases.wick_1d("D1")
    over = dict(cfg, get_variance_1d={"D1": var1}, xi_1d={"D1": xi1}, xi_wick=None)
    helpers.configure(xcf, data, num, cases.ang_max_for(cosmo, cfg, z_min, z_min2), objs=objs,
                      **over)
    for k, v in over.items():
        setattr(xcf, k, v)
    hps = sorted(data)
    xcf.fill_neighs(hps)
    np.random.seed(hps[0])  # picca_xwick.py:36
    res = xcf.compute_wick_terms(hps)
    pre = "xwick_%s_" % name
    assert (res[2], res[3]) == tuple(gold[pre + "counts"])
    assert np.array_equal(res[1], gold[pre + "num_pairs_wick"]) and res[1].sum() > 5000
    np.testing.assert_allclose(res[0], gold[pre + "weights_wick"], rtol=1e-9)
    for k in (1, 2, 3, 4):
        close(res[3 + k], gold[pre + "t%d" % k], "%s t%d" % (name, k))
    assert not res[8].any() and not res[9].any()


def test_xwick_zero_weight_pixel_raises_like_numba():
    """xcf.py:1316 divides by the pixel weight: the reference (Numba, Python error model) raises
    ZeroDivisionError when a zero-weight pixel is in range; so does the B200 path."""
    from picca_b200 import xcf
    cfg = cases.XWICK_CASES["default"]
    data, num, z_min, cosmo = cases.dmat_forests()   # 2 % zero-weight pixels
    cases.set_fname(data, "D1")
    objs, z_min2 = cases.quasars(cosmo)
    var1, xi1 = cases.wick_1d("D1")
    over = dict(cfg, get_variance_1d={"D1": var1}, xi_1d={"D1": xi1}, xi_wick=None, reject=0.)
    helpers.configure(xcf, data, num, cases.ang_max_for(cosmo, cfg, z_min, z_min2), objs=objs,
                      **over)
    for k, v in over.items():
        setattr(xcf, k, v)
    hps = sorted(data)
    xcf.fill_neighs(hps)
    with pytest.raises(ZeroDivisionError):
        xcf.compute_wick_terms(hps)


def test_wick_against_oracle_longer_forests_and_production_binning():
    """other inputs than the goldens: longer forests (several 32-lane steps per row, many bins per
    row), 50 x 50 bins out to 200 Mpc/h; oracle C restatement on the same draw."""
    from oracle import cf as ocf
    from picca_b200 import cf, synth
    data, num, z_min, _, cosmo = helpers.small_sample(n=40, seed=91, max_pix=90, side_deg=2.5)
    cases.set_fname(data, "D1")
    var1, xi1 = cases.wick_1d("D1")
    over = dict(r_par_max=200., r_trans_max=200., num_bins_r_par=50, num_bins_r_trans=50,
                reject=0.9, max_diagram=3, get_variance_1d={"D1": var1}, xi_1d={"D1": xi1})
    ang_max = synth.compute_ang_max(cosmo, 200., z_min)
    hps = sorted(data)
    results = []
    for mod in (ocf, cf):
        helpers.configure(mod, data, num, ang_max, **over)
        for k, v in over.items():
            setattr(mod, k, v)
        mod.fill_neighs(hps)
        np.random.seed(7)
        results.append(mod.compute_wick_terms(hps))
    want, got = results
    assert (want[2], want[3]) == (got[2], got[3]) and want[3] >= 2
    assert np.array_equal(want[1], got[1]) and want[1].sum() > 10000
    np.testing.assert_allclose(got[0], want[0], rtol=1e-9)
    for k in (4, 5, 6):
        close(got[k], want[k], "t%d" % (k - 3))


def test_wick_rejects_what_is_not_built():
    from picca_b200 import cf
    data = setup_wick(cf, cases.WICK_CASES["default"])
    hps = sorted(data)
    cf.fill_neighs(hps)
    cf.max_diagram = 4
    with pytest.raises(NotImplementedError):
        cf.compute_wick_terms(hps)
    cf.max_diagram = 3
    cf.xi_1d = {"D1": lambda dll: np.exp(-dll)}   # not a nearest-neighbour table
    with pytest.raises(NotImplementedError):
        np.random.seed(1)
        cf.compute_wick_terms(hps)
