"""CPU: the oracle (C restatement + NumPy host logic) against the golden vectors produced by the
LIVE reference (tests/golden/make_golden.py).  This is what pins the oracle wherever the reference
itself cannot run (the GPU box)."""
import os

import numpy as np
import pytest

from tests import helpers
from tests.golden import cases

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(tag):
    return np.load(os.path.join(GOLD, "golden_%s.npz" % tag))


def neighbour_ids(data, hps):
    counts, ids = [], []
    for hp in hps:
        for d in data[hp]:
            counts.append(len(d.neighbours))
            ids.extend(int(o.thingid) for o in d.neighbours)
    return np.array(counts, dtype=np.int64), np.array(ids, dtype=np.int64)


def check_rows(rows, gold, rtol):
    rows = np.stack(rows)
    assert np.array_equal(rows[:, 5].view(np.int64), gold[:, 5].view(np.int64))
    for k in range(5):
        np.testing.assert_allclose(rows[:, k], gold[:, k], rtol=rtol, atol=1e-300)


def setup_cf(mod, cfg, dmat=False):
    cfg = dict(cfg)
    second = cfg.pop("second", False)
    src = cases.dmat_forests if dmat else cases.forests
    data, num, z_min, cosmo = src()
    over, z_min2 = dict(cfg), None
    if second:
        data2, num2, z_min2, _ = src(second=True)
        over["data2"], over["num_data2"] = data2, num2
    helpers.configure(mod, data, num, cases.ang_max_for(cosmo, cfg, z_min, z_min2), **over)
    return data


@pytest.mark.parametrize("name", sorted(cases.CF_CASES))
def test_cf(name):
    from oracle import cf as ocf
    gold = load("cf")
    data = setup_cf(ocf, cases.CF_CASES[name])
    rows, counts, ids = [], [], []
    for hp in sorted(data):
        ocf.fill_neighs([hp])
        c, i = neighbour_ids(data, [hp])
        counts.append(c)
        ids.append(i)
        res = ocf.compute_xi([hp])
        rows.append(np.stack([np.asarray(r, dtype=np.float64) for r in res[:5]] +
                             [np.asarray(res[5]).view(np.float64)]))
    assert np.array_equal(np.concatenate(counts), gold["cf_%s_nbcount" % name])
    assert np.array_equal(np.concatenate(ids), gold["cf_%s_nbid" % name])
    check_rows(rows, gold["cf_%s" % name], rtol=1e-12)


def check8(res, gold, prefix, rtol):
    names = ("weights_dmat", "dmat", "r_par_eff", "r_trans_eff", "z_eff", "weight_eff")
    assert [int(res[6]), int(res[7])] == list(gold[prefix + "counts"])
    for k, n in enumerate(names):
        want = gold[prefix + n]
        scale = np.abs(want).max()
        np.testing.assert_allclose(res[k], want, rtol=rtol, atol=rtol * scale)


@pytest.mark.parametrize("name", sorted(cases.DMAT_CASES))
def test_dmat(name):
    from oracle import cf as ocf
    gold = load("dmat")
    data = setup_cf(ocf, cases.DMAT_CASES[name], dmat=True)
    hps = sorted(data)
    ocf.fill_neighs(hps)
    np.random.seed(hps[0])
    res = ocf.compute_dmat(hps)
    check8(res, gold, "dmat_%s_" % name, rtol=1e-12)


@pytest.mark.parametrize("name", sorted(cases.XCF_CASES))
def test_xcf(name):
    from oracle import xcf as oxcf
    gold = load("xcf")
    cfg = cases.XCF_CASES[name]
    data, num, z_min, cosmo = cases.forests()
    objs, z_min2 = cases.quasars(cosmo)
    helpers.configure(oxcf, data, num, cases.ang_max_for(cosmo, cfg, z_min, z_min2), objs=objs,
                      **cfg)
    rows, counts, ids = [], [], []
    for hp in sorted(data):
        oxcf.fill_neighs([hp])
        c, i = neighbour_ids(data, [hp])
        counts.append(c)
        ids.append(i)
        res = oxcf.compute_xi([hp])
        rows.append(np.stack([np.asarray(r, dtype=np.float64) for r in res[:5]] +
                             [np.asarray(res[5]).view(np.float64)]))
    assert np.array_equal(np.concatenate(counts), gold["xcf_%s_nbcount" % name])
    assert np.array_equal(np.concatenate(ids), gold["xcf_%s_nbid" % name])
    check_rows(rows, gold["xcf_%s" % name], rtol=1e-12)


@pytest.mark.parametrize("name", sorted(cases.XDMAT_CASES))
def test_xdmat(name):
    from oracle import xcf as oxcf
    gold = load("xdmat")
    cfg = cases.XDMAT_CASES[name]
    data, num, z_min, cosmo = cases.dmat_forests()
    objs, z_min2 = cases.quasars(cosmo)
    helpers.configure(oxcf, data, num, cases.ang_max_for(cosmo, cfg, z_min, z_min2), objs=objs,
                      **cfg)
    hps = sorted(data)
    oxcf.fill_neighs(hps)
    np.random.seed(hps[0])
    res = oxcf.compute_dmat(hps)
    check8(res, gold, "xdmat_%s_" % name, rtol=1e-12)
