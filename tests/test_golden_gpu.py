"""GPU parity against the golden vectors of the LIVE reference (tests/golden/): picca_b200.cf on
the CUDA path must reproduce the reference's neighbour lists exactly, num_pairs bit for bit and
the fp64 sums within 1e-9 relative (north_star tolerance)."""
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


@pytest.mark.parametrize("host_angles", [False, True])
@pytest.mark.parametrize("name", sorted(cases.CF_CASES))
def test_cf_matches_reference_golden(name, host_angles, monkeypatch):
    from picca_b200 import _corr, cf
    monkeypatch.setattr(_corr, "HOST_ANGLES", host_angles)
    gold = np.load(os.path.join(GOLD, "golden_cf.npz"))
    cfg = dict(cases.CF_CASES[name])
    second = cfg.pop("second", False)
    plates = cfg.pop("plates", False)
    data, num, z_min, cosmo = cases.forests(plates=plates)
    over, z_min2 = dict(cfg), None
    if second:
        data2, num2, z_min2, _ = cases.forests(second=True, plates=plates)
        over["data2"], over["num_data2"] = data2, num2
    helpers.configure(cf, data, num, cases.ang_max_for(cosmo, cfg, z_min, z_min2), **over)
    want = gold["cf_%s" % name]
    counts, ids = [], []
    for k, hp in enumerate(sorted(data)):
        cf.fill_neighs([hp])
        c, i = neighbour_ids(data, [hp])
        counts.append(c)
        ids.append(i)
        got = cf.compute_xi([hp])
        ref = [want[k, f] for f in range(5)] + [want[k, 5].view(np.int64)]
        helpers.assert_xi_close(got, ref, tag="%s hp %d" % (name, hp))
    assert np.array_equal(np.concatenate(counts), gold["cf_%s_nbcount" % name])
    assert np.array_equal(np.concatenate(ids), gold["cf_%s_nbid" % name])


def test_batch_equals_per_healpix_calls():
    from picca_b200 import cf
    cfg = cases.CF_CASES["default"]
    data, num, z_min, cosmo = cases.forests()
    helpers.configure(cf, data, num, cases.ang_max_for(cosmo, cfg, z_min), **cfg)
    hps = sorted(data)
    cf.fill_neighs(hps)
    block = cf.compute_xi_batch(hps)
    for k, hp in enumerate(hps):
        cf.fill_neighs([hp])
        one = cf.compute_xi([hp])
        got = [block[k, f] for f in range(5)] + [block[k, 5].view(np.int64)]
        helpers.assert_xi_close(got, one, tag="batch hp %d" % hp)
    # compute_xi over several pixels sums them into ONE histogram (cf.py:154-161)
    cf.fill_neighs(hps[:3])
    merged = cf.compute_xi(hps[:3])
    assert int(merged[5].sum()) == int(block[:3, 5].view(np.int64).sum())
