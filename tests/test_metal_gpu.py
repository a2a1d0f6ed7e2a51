"""GPU parity of the metal distortion matrix: picca_b200.cf.compute_metal_dmat (pb2_metal_dmat_auto
through the C ABI) against the live reference's golden vectors (tests/golden/golden_metal.npz)
and against the oracle restatement at production binning.  NPALL / NPUSED exact; which data bins
receive weight must match exactly (bit-exact bins); sums within 1e-9 relative with an absolute
floor of 1e-12 x the largest entry (atomics re-associate the sums)."""
import os

import numpy as np
import pytest

from tests import helpers
from tests.golden import cases
from tests.test_dmat_gpu import check8

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def setup_metal(mod, cfg, forests=cases.forests):
    cfg = dict(cfg)
    second = cfg.pop("second", False)
    pair = cfg.pop("pair")
    data, num, z_min, cosmo = forests()
    over, z_min2 = dict(cfg, alpha_abs=dict(cases.ALPHA_ABS), cosmo=cosmo), None
    if second:
        data2, num2, z_min2, _ = forests(second=True)
        over["data2"], over["num_data2"] = data2, num2
    helpers.configure(mod, data, num, cases.ang_max_for(cosmo, cfg, z_min, z_min2), **over)
    for k, v in over.items():
        setattr(mod, k, v)
    return data, pair


@pytest.mark.parametrize("name", sorted(cases.METAL_CASES))
def test_metal_dmat_matches_reference_golden(name):
    from picca_b200 import cf
    gold = np.load(os.path.join(GOLD, "golden_metal.npz"))
    data, pair = setup_metal(cf, cases.METAL_CASES[name])
    hps = sorted(data)
    cf.fill_neighs(hps)
    np.random.seed(hps[0])  # picca_metal_dmat.py:48
    res = cf.compute_metal_dmat(hps, abs_igm1=pair[0], abs_igm2=pair[1])
    check8(res, gold, "metal_%s_" % name)
    want = gold["metal_%s_dmat" % name]
    assert np.array_equal(res[1] != 0, want != 0)  # the same (data bin, model bin) cells are hit
    assert all(d.neighbours is None for hp in hps for d in data[hp])  # cf.py:1219


def test_metal_dmat_chunks_sum_like_the_script():
    """picca_metal_dmat.py sums the per-chunk results of its workers (:494-501 pattern): two
    chunks with their own seeds equal the oracle on the same chunks."""
    from oracle import cf as ocf
    from picca_b200 import cf
    cfg = cases.METAL_CASES["lya_si3"]
    total = {}
    for mod in (cf, ocf):
        data, pair = setup_metal(mod, cfg)
        hps = sorted(data)
        chunks = [hps[0::2], hps[1::2]]
        acc = None
        for chunk in chunks:
            mod.fill_neighs(chunk)
            np.random.seed(chunk[0])
            res = mod.compute_metal_dmat(chunk, abs_igm1=pair[0], abs_igm2=pair[1])
            acc = list(res) if acc is None else [a + b for a, b in zip(acc, res)]
        total[mod.__name__] = acc
    got, want = total[cf.__name__], total[ocf.__name__]
    assert (got[6], got[7]) == (want[6], want[7])
    for a, b in zip(got[:6], want[:6]):
        np.testing.assert_allclose(a, b, rtol=1e-9, atol=1e-12 * np.abs(b).max())


def test_metal_dmat_production_binning_matches_oracle():
    """np = nt = 50, 200 Mpc/h (the binning of BASELINE config 4) on the small sample."""
    from oracle import cf as ocf
    from picca_b200 import cf
    cfg = dict(r_par_max=200., r_trans_max=200., num_bins_r_par=50, num_bins_r_trans=50,
               num_model_bins_r_par=50, num_model_bins_r_trans=50, reject=0.95,
               pair=("LYA", "SiIII(1207)"))
    out = []
    for mod in (cf, ocf):
        data, pair = setup_metal(mod, cfg, forests=cases.dmat_forests)
        hps = sorted(data)
        mod.fill_neighs(hps)
        np.random.seed(hps[0])
        out.append(mod.compute_metal_dmat(hps, abs_igm1=pair[0], abs_igm2=pair[1]))
    got, want = out
    assert (got[6], got[7]) == (want[6], want[7]) and want[7] > 0
    assert np.array_equal(got[1] != 0, want[1] != 0)
    for a, b in zip(got[:6], want[:6]):
        np.testing.assert_allclose(a, b, rtol=1e-9, atol=1e-12 * np.abs(b).max())


@pytest.mark.parametrize("name", sorted(cases.XMETAL_CASES))
def test_xcf_metal_dmat_matches_reference_golden(name):
    """xcf.compute_metal_dmat (xcf.py:677-835) incl. the forests skipped before the --rej draw."""
    from picca_b200 import cf, xcf
    gold = np.load(os.path.join(GOLD, "golden_xmetal.npz"))
    cfg = dict(cases.XMETAL_CASES[name])
    abs_igm = cfg.pop("abs_igm")
    cf.absorber_igm.update(cases.EXTRA_ABSORBERS)
    data, num, z_min, cosmo = cases.forests()
    objs, z_min2 = cases.quasars(cosmo)
    over = dict(cfg, alpha_abs=dict(cases.ALPHA_ABS), cosmo=cosmo)
    helpers.configure(xcf, data, num, cases.ang_max_for(cosmo, cfg, z_min, z_min2), objs=objs,
                      **over)
    for k, v in over.items():
        setattr(xcf, k, v)
    hps = sorted(data)
    xcf.fill_neighs(hps)
    np.random.seed(hps[0])
    res = xcf.compute_metal_dmat(hps, abs_igm=abs_igm)
    check8(res, gold, "xmetal_%s_" % name)
    assert np.array_equal(res[1] != 0, gold["xmetal_%s_dmat" % name] != 0)
    kept = sum(d.neighbours is not None for hp in hps for d in data[hp])
    assert kept == int(gold["xmetal_%s_skipped" % name][0])  # xcf.py:741-742
