"""GPU parity: picca_b200.xcf (CUDA) against the live reference's golden vectors and the oracle."""
import os

import numpy as np
import pytest

from tests import helpers
from tests.golden import cases

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def neighbour_ids(data, hps):
    counts, ids = [], []
    for hp in hps:
        for d in data[hp]:
            counts.append(len(d.neighbours))
            ids.extend(int(o.thingid) for o in d.neighbours)
    return np.array(counts, dtype=np.int64), np.array(ids, dtype=np.int64)


@pytest.mark.parametrize("variant", [0, 1, 2, 4])  # prefix-sum kernel (transposed reductions), validation, lane = object kernel, prefix-sum kernel with per-lane reductions
@pytest.mark.parametrize("name", sorted(cases.XCF_CASES))
def test_xcf_matches_reference_golden(name, variant):
    from picca_b200 import xcf
    gold = np.load(os.path.join(GOLD, "golden_xcf.npz"))
    cfg = cases.XCF_CASES[name]
    data, num, z_min, cosmo = cases.forests()
    objs, z_min2 = cases.quasars(cosmo)
    helpers.configure(xcf, data, num, cases.ang_max_for(cosmo, cfg, z_min, z_min2), objs=objs,
                      **cfg)
    xcf._XI_VARIANT = variant
    want = gold["xcf_%s" % name]
    counts, ids = [], []
    for k, hp in enumerate(sorted(data)):
        xcf.fill_neighs([hp])
        c, i = neighbour_ids(data, [hp])
        counts.append(c)
        ids.append(i)
        got = xcf.compute_xi([hp])
        ref = [want[k, f] for f in range(5)] + [want[k, 5].view(np.int64)]
        helpers.assert_xi_close(got, ref, tag="%s hp %d" % (name, hp))
    xcf._XI_VARIANT = 0
    assert np.array_equal(np.concatenate(counts), gold["xcf_%s_nbcount" % name])
    assert np.array_equal(np.concatenate(ids), gold["xcf_%s_nbid" % name])


def test_xcf_production_binning_matches_oracle():
    """np=100, nt=50, rp in [-200,200] (picca_xcf.py defaults, :83-112) on a denser patch."""
    from oracle import xcf as oxcf
    from picca_b200 import synth, xcf
    data, num, z_min, _, cosmo = helpers.small_sample(n=150, seed=41, max_pix=200, side_deg=5.)
    objs, z_min2 = synth.make_quasars(300, seed=43, nside=16, ra_deg=(10., 15.),
                                      dec_deg=(5., 10.), z_range=(1.9, 3.3), cosmo=cosmo)
    cfg = dict(r_par_max=200., r_par_min=-200., r_trans_max=200., num_bins_r_par=100,
               num_bins_r_trans=50, alpha_obj=1.44)
    ang_max = synth.compute_ang_max(cosmo, 200., z_min, z_min2)
    helpers.configure(oxcf, data, num, ang_max, objs=objs, **cfg)
    helpers.configure(xcf, data, num, ang_max, objs=objs, **cfg)
    hps = sorted(data)
    xcf.fill_neighs(hps)
    block = xcf.compute_xi_batch(hps)
    total = 0
    for k, hp in enumerate(hps):
        oxcf.fill_neighs([hp])
        want = oxcf.compute_xi([hp])
        got = [block[k, f] for f in range(5)] + [block[k, 5].view(np.int64)]
        helpers.assert_xi_close(got, want, tag="hp %d" % hp)
        total += int(want[5].sum())
    assert total > 10**5
