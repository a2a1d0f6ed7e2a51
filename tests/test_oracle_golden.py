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
    plates = cfg.pop("plates", False)
    src = cases.dmat_forests if dmat else cases.forests
    data, num, z_min, cosmo = src()
    if plates:
        cases.share_plates(data)
    over, z_min2 = dict(cfg), None
    if second:
        data2, num2, z_min2, _ = src(second=True)
        if plates:
            cases.share_plates(data2)
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


# ---- covariance step (oracle/export.py against the live reference's outputs)
def _cov_close(got, want, rtol):
    """covariance entries compared on the scale of the two variances (an off-diagonal entry is a
    cancelling sum, its own magnitude is not a scale)"""
    sd = np.sqrt(np.abs(np.diagonal(want)))
    scale = np.maximum(sd[:, None] * sd[None, :], np.abs(want))
    assert got.shape == want.shape
    assert np.all(np.abs(got - want) <= rtol * scale + 1e-300), \
        "max scaled error %.3e" % np.max(np.abs(got - want) / np.maximum(scale, 1e-300))


@pytest.mark.parametrize("name", ["small", "ragged", "per_r_par", "xcf_like", "empty_bin"])
def test_export_covariance(name):
    from oracle import export as oexp
    from tests.golden import cases_export
    gold = load("export")
    cfg = cases_export.CASES[name]
    xi, we, rp, rt = cases_export.inputs(cfg)
    cov = oexp.compute_cov(xi, we)
    _cov_close(cov, gold["%s_cov" % name], 1e-12)
    smooth = oexp.smooth_cov(xi, we, rp, rt, delta_r_trans=cfg["delta_r_trans"],
                             delta_r_par=cfg["delta_r_par"], covariance=gold["%s_cov" % name],
                             per_r_par=cfg.get("per_r_par", False))
    _cov_close(smooth, gold["%s_smooth" % name], 1e-12)
    if name == "empty_bin":  # zero variance: the reference returns the covariance unsmoothed
        assert np.array_equal(smooth, gold["%s_cov" % name])
    with np.errstate(all="ignore"):
        boot = oexp.compute_cov_boot(xi, we, nboots=cases_export.NBOOTS,
                                     seed=cases_export.BOOT_SEED)
    want = gold["%s_boot" % name]
    assert np.array_equal(np.isnan(boot), np.isnan(want))  # an empty bin: 0/0 in every realisation
    ok = ~np.isnan(want)
    _cov_close(np.where(ok, boot, 0.), np.where(ok, want, 0.), 1e-10)


def test_export_covariance_reference_fixture():
    """cf.fits.gz -> exported_cf.fits.gz, the reference's own golden pair (test_3_cor.py:443-456,
    compared there at rtol 1e-5): sub-sample covariance + smoothing with the script's bin widths
    (picca_export.py:271-293)."""
    from oracle import export as oexp
    gold = load("export")
    n_p, n_t, rp_min, rp_max, rt_max = gold["fixture_bins"]
    cov = oexp.compute_cov(gold["fixture_da"], gold["fixture_we"])
    smooth = oexp.smooth_cov(None, None, gold["fixture_rp"], gold["fixture_rt"],
                             delta_r_trans=(rt_max - 0.) / n_t,
                             delta_r_par=(rp_max - rp_min) / n_p, covariance=cov)
    np.testing.assert_allclose(smooth, gold["fixture_co"], rtol=1e-5, atol=1e-8 * np.abs(
        gold["fixture_co"]).max())


# ---- delta loader (oracle/io.py against the live reference's io.read_deltas outputs)
IO_RUNS = {"fixture": None, "imagefixture": None, "image": {}, "imageblind": {}, "sdss": {}, "desi": {}, "blind": {},
           "sdss_noproject": dict(no_project=True), "sdss_max30": dict(max_num_spec=30),
           "sdss_zcut": dict(z_min_qso=2.4, z_max_qso=3.0),
        "lin_rebin2": dict(rebin_factor=2), "image_rebin3": dict(rebin_factor=3),
        "imagefixture_rebin3": dict(rebin_factor=3)}


def io_inputs(tag, tmp_path):
    from tests.golden import cases_io
    fx = os.path.join(GOLD, "fixtures")
    if tag == "fixture":
        return os.path.join(fx, "delta-272.fits.gz"), os.path.join(fx, "delta_attributes.fits.gz")
    if tag.startswith("imagefixture"):
        return (os.path.join(fx, "image-delta-50.fits.gz"),
                os.path.join(fx, "delta_attributes.fits.gz"))
    if tag.split("_")[0] in cases_io.IMAGE_CASES:
        return cases_io.write_image_case(str(tmp_path), tag.split("_")[0])
    return cases_io.write_case(str(tmp_path), tag.split("_")[0])


@pytest.mark.parametrize("tag", sorted(IO_RUNS))
def test_read_deltas(tag, tmp_path):
    from oracle import io as oio
    from tests.golden import cases_io
    gold = load("io")
    in_dir, attr = io_inputs(tag, tmp_path)
    tables = (gold["cosmo_z"], gold["cosmo_r_comov"], gold["cosmo_dist_m"])
    data, num, z_min, z_max = oio.read_deltas(in_dir, tables=tables, delta_attributes=attr,
                                              **dict(cases_io.READ_KW, **(IO_RUNS[tag] or {})))
    flat = cases_io.flatten(data)
    assert [num, z_min, z_max] == list(gold["%s_summary" % tag])
    for k, v in flat.items():
        want = gold["%s_%s" % (tag, k)]
        if v.dtype.kind == "f" and k in ("weights", "delta"):
            # NumPy's pairwise sums / power may differ between SIMD back-ends of two hosts
            np.testing.assert_allclose(v, want, rtol=1e-12, atol=1e-14, err_msg=k)
        else:
            assert np.array_equal(v, want), k


# ---- metal distortion matrix (oracle compute_metal_dmat against the live reference's outputs)
def setup_metal(mod, cfg):
    cfg = dict(cfg)
    second = cfg.pop("second", False)
    pair = cfg.pop("pair")
    data, num, z_min, cosmo = cases.forests()
    over, z_min2 = dict(cfg, alpha_abs=dict(cases.ALPHA_ABS), cosmo=cosmo), None
    if second:
        data2, num2, z_min2, _ = cases.forests(second=True)
        over["data2"], over["num_data2"] = data2, num2
    helpers.configure(mod, data, num, cases.ang_max_for(cosmo, cfg, z_min, z_min2), **over)
    for k, v in over.items():
        setattr(mod, k, v)
    return data, pair


@pytest.mark.parametrize("name", sorted(cases.METAL_CASES))
def test_metal_dmat(name):
    from oracle import cf as ocf
    gold = load("metal")
    data, pair = setup_metal(ocf, cases.METAL_CASES[name])
    hps = sorted(data)
    ocf.fill_neighs(hps)
    np.random.seed(hps[0])
    res = ocf.compute_metal_dmat(hps, abs_igm1=pair[0], abs_igm2=pair[1])
    check8(res, gold, "metal_%s_" % name, rtol=1e-12)


def setup_xmetal(mod, cfg, absorbers):
    cfg = dict(cfg)
    abs_igm = cfg.pop("abs_igm")
    absorbers.update(cases.EXTRA_ABSORBERS)
    data, num, z_min, cosmo = cases.forests()
    objs, z_min2 = cases.quasars(cosmo)
    over = dict(cfg, alpha_abs=dict(cases.ALPHA_ABS), cosmo=cosmo)
    helpers.configure(mod, data, num, cases.ang_max_for(cosmo, cfg, z_min, z_min2), objs=objs,
                      **over)
    for k, v in over.items():
        setattr(mod, k, v)
    return data, abs_igm


@pytest.mark.parametrize("name", sorted(cases.XMETAL_CASES))
def test_xmetal_dmat(name):
    from oracle import xcf as oxcf
    gold = load("xmetal")
    data, abs_igm = setup_xmetal(oxcf, cases.XMETAL_CASES[name], oxcf.absorber_igm)
    hps = sorted(data)
    oxcf.fill_neighs(hps)
    np.random.seed(hps[0])
    res = oxcf.compute_metal_dmat(hps, abs_igm=abs_igm)
    check8(res, gold, "xmetal_%s_" % name, rtol=1e-12)
    kept = sum(d.neighbours is not None for hp in hps for d in data[hp])
    assert kept == int(gold["xmetal_%s_skipped" % name][0])  # xcf.py:741-742


# ---- object x object correlation (oracle/co.py against the live reference's outputs)
def setup_co(mod, cfg):
    from picca_b200 import synth
    cfg = dict(cfg)
    second = cfg.pop("second", False)
    cosmo = synth.FlatLCDM()
    objs, z_min = cases.quasars(cosmo)
    for k, v in dict(dict(x_correlation=False), **cfg).items():
        setattr(mod, k, v)
    mod.objs, mod.objs2, z_min2 = objs, None, None
    if second:
        mod.objs2, z_min2 = cases.quasars2(cosmo)
    mod.ang_max = cases.ang_max_for(cosmo, cfg, z_min, z_min2)
    mod.nside = 16
    mod.num_data = sum(len(v) for v in objs.values())
    mod.lock, mod.counter = helpers.DummyLock(), helpers.DummyCounter()
    return objs


@pytest.mark.parametrize("name", sorted(cases.CO_CASES))
def test_co(name):
    from oracle import co as oco
    gold = load("co")["co_%s" % name]
    objs = setup_co(oco, cases.CO_CASES[name])
    rows = []
    for hp in sorted(objs):
        oco.fill_neighs([hp])
        res = oco.compute_xi([hp])
        rows.append(np.stack([np.asarray(r, dtype=np.float64) for r in res[:4]] +
                             [np.asarray(res[4], dtype=np.int64).view(np.float64)]))
    rows = np.stack(rows)
    assert np.array_equal(rows[:, 4].view(np.int64), gold[:, 4].view(np.int64))
    for k in range(4):
        np.testing.assert_allclose(rows[:, k], gold[:, k], rtol=1e-12, atol=1e-300)
