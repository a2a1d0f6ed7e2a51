"""CPU, this container only: pin the oracle against (1) the LIVE reference on the reference's own
bundled fixtures, bit for bit, and (2) the reference's golden FITS files, by running the
reference's UNMODIFIED CLI scripts on top of the oracle's module doubles with the exact command
lines of reference py/picca/tests/test_3_cor.py (:229-257, :318-349, :636-698).
Skipped where /root/reference does not exist (the GPU box)."""
import numpy as np
import pytest

from tests.refharness import shims

pytestmark = [pytest.mark.reference,
              pytest.mark.skipif(not shims.reference_available(), reason="needs /root/reference")]

DATA = shims.REFERENCE_DATA


@pytest.fixture(scope="module")
def fixture_data():
    from tests.refharness import load
    data, num_data, z_min, z_max, cosmo = load.load_deltas()
    objs, z_min2 = load.load_objects(cosmo)
    return data, num_data, z_min, cosmo, objs, z_min2


def _setup(mod, fx, utils, load, cross_obj=False, **over):
    data, num_data, z_min, cosmo, objs, z_min2 = fx
    cfg = dict(r_par_max=60., r_trans_max=60., r_par_min=0., num_bins_r_par=15,
               num_bins_r_trans=15, num_model_bins_r_par=15, num_model_bins_r_trans=15, nside=16,
               z_ref=2.25, alpha=2.9, alpha2=2.9, reject=0.99)
    cfg.update(over)
    from tests import helpers
    for k, v in dict(helpers.CF_DEFAULTS, **cfg).items():  # no state leaks from earlier tests
        setattr(mod, k, v)
    mod.data, mod.num_data = data, num_data
    if cross_obj:
        mod.objs = objs
        mod.ang_max = utils.compute_ang_max(cosmo, mod.r_trans_max, z_min, z_min2)
    else:
        mod.ang_max = utils.compute_ang_max(cosmo, mod.r_trans_max, z_min)
    mod.lock, mod.counter = load.DummyLock(), load.DummyCounter()


def test_cf_bit_exact_on_bundled_fixtures(fixture_data):
    from tests.refharness import load
    from oracle import cf as ocf
    cf, _, _, _, utils = load.reference_modules()
    cf.userprint = lambda *a, **k: None
    data = fixture_data[0]
    over = dict(remove_same_half_plate_close_pairs=True)
    _setup(cf, fixture_data, utils, load, **over)
    _setup(ocf, fixture_data, utils, load, **over)
    total = 0
    for hp in sorted(data)[:8]:
        cf.fill_neighs([hp])
        want_n = [[d2.thingid for d2 in d.neighbours] for d in data[hp]]
        want = cf.compute_xi([hp])
        ocf.fill_neighs([hp])
        assert [[d2.thingid for d2 in d.neighbours] for d in data[hp]] == want_n
        got = ocf.compute_xi([hp])
        for a, b in zip(want, got):
            assert np.array_equal(a, b)
        total += int(got[5].sum())
    assert total > 10**6


@pytest.mark.parametrize("evol", [False, True])
def test_dmat_bit_exact_on_bundled_fixtures(fixture_data, evol):
    from tests.refharness import load
    from oracle import cf as ocf
    cf, _, _, _, utils = load.reference_modules()
    cf.userprint = lambda *a, **k: None
    hps = sorted(fixture_data[0])
    over = dict(remove_same_half_plate_close_pairs=True,
                redshift_evolution_in_distortion_matrix=evol)
    _setup(cf, fixture_data, utils, load, **over)
    _setup(ocf, fixture_data, utils, load, **over)
    cf.fill_neighs(hps)
    np.random.seed(hps[0])
    want = cf.compute_dmat(hps)
    ocf.fill_neighs(hps)
    np.random.seed(hps[0])
    got = ocf.compute_dmat(hps)
    assert (want[6], want[7]) == (11845, 121) == (got[6], got[7])  # golden NPALL / NPUSED
    for a, b in zip(want[:6], got[:6]):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("abs1,abs2,cross", [("LYA", "SiIII(1207)", False),
                                             ("SiIII(1207)", "SiIII(1207)", False),
                                             ("LYA", "SiII(1190)", True)])
def test_metal_dmat_bit_exact_on_bundled_fixtures(fixture_data, abs1, abs2, cross):
    """oracle compute_metal_dmat against the live cf.compute_metal_dmat (cf.py:890-1232) with the
    flags of test_3_cor.py:351-381 (auto) and :526-560 (cross: --in-dir2, --unfold-cf,
    --remove-same-half-plate-close-pairs)."""
    from tests.refharness import load
    from oracle import cf as ocf
    cf, _, _, _, utils = load.reference_modules()
    cf.userprint = lambda *a, **k: None
    hps = sorted(fixture_data[0])
    over = dict(alpha_abs={"LYA": 2.9, "SiIII(1207)": 1., "SiII(1190)": 1.},
                cosmo=fixture_data[3])
    if cross:
        over.update(data2=fixture_data[0], num_data2=fixture_data[1], x_correlation=True,
                    r_par_min=-60., num_bins_r_par=30, num_model_bins_r_par=30,
                    remove_same_half_plate_close_pairs=True, lambda_abs="LYA", lambda_abs2="LYA")
    results = []
    for mod in (cf, ocf):
        _setup(mod, fixture_data, utils, load, **over)
        for k, v in over.items():
            setattr(mod, k, v)
        mod.fill_neighs(hps)
        np.random.seed(hps[0])
        results.append(mod.compute_metal_dmat(hps, abs_igm1=abs1, abs_igm2=abs2))
    want, got = results
    assert (want[6], want[7]) == (got[6], got[7]) and want[7] > 50
    assert want[1].sum() > 0
    for a, b in zip(want[:6], got[:6]):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("abs_igm", ["SiIII(1207)", "SiII(1190)"])
def test_xcf_metal_dmat_bit_exact_on_bundled_fixtures(fixture_data, abs_igm):
    """oracle xcf.compute_metal_dmat against the live one (xcf.py:677-835)."""
    from tests.refharness import load
    from oracle import xcf as oxcf
    _, xcf, _, _, utils = load.reference_modules()
    xcf.userprint = lambda *a, **k: None
    hps = sorted(fixture_data[0])
    over = dict(r_par_min=-60., num_bins_r_par=30, num_model_bins_r_par=30, alpha_obj=1.,
                reject=0.9, alpha_abs={"LYA": 2.9, "SiIII(1207)": 1., "SiII(1190)": 1.},
                cosmo=fixture_data[3])
    results = []
    for mod in (xcf, oxcf):
        _setup(mod, fixture_data, utils, load, cross_obj=True, **over)
        for k, v in over.items():
            setattr(mod, k, v)
        mod.fill_neighs(hps)
        np.random.seed(hps[0])
        results.append(mod.compute_metal_dmat(hps, abs_igm=abs_igm))
    want, got = results
    assert (want[6], want[7]) == (got[6], got[7]) and want[7] > 50
    assert want[1].sum() > 0
    for a, b in zip(want[:6], got[:6]):
        assert np.array_equal(a, b)


def test_xcf_and_xdmat_bit_exact_on_bundled_fixtures(fixture_data):
    from tests.refharness import load
    from oracle import xcf as oxcf
    _, xcf, _, _, utils = load.reference_modules()
    xcf.userprint = lambda *a, **k: None
    data = fixture_data[0]
    hps = sorted(data)
    over = dict(r_par_min=-60., num_bins_r_par=30, num_model_bins_r_par=30, alpha_obj=1.,
                redshift_evolution_in_distortion_matrix=False)
    _setup(xcf, fixture_data, utils, load, cross_obj=True, **over)
    _setup(oxcf, fixture_data, utils, load, cross_obj=True, **over)
    total = 0
    for hp in hps:
        xcf.fill_neighs([hp])
        want = xcf.compute_xi([hp])
        oxcf.fill_neighs([hp])
        got = oxcf.compute_xi([hp])
        for a, b in zip(want, got):
            assert np.array_equal(a, b)
        total += int(got[5].sum())
    assert total == 353334  # NB total of the reference's golden xcf.fits.gz
    xcf.fill_neighs(hps)
    np.random.seed(hps[0])
    want = xcf.compute_dmat(hps)
    oxcf.fill_neighs(hps)
    np.random.seed(hps[0])
    got = oxcf.compute_dmat(hps)
    assert (want[6], want[7]) == (1202, 102) == (got[6], got[7])
    for a, b in zip(want[:6], got[:6]):
        assert np.array_equal(a, b)


# ---------------------------------------------------------------------------------------------
# unmodified scripts on top of the oracle doubles vs the reference's golden FITS
# ---------------------------------------------------------------------------------------------
def _compare_fits(path_got, path_want):
    """reference tests/test_helpers.py:45-112: same columns, array_equal else allclose(1e-5,1e-8)"""
    from tests.refharness import minifits
    got, want = minifits.FITS(path_got), minifits.FITS(path_want)
    assert len(got) == len(want)
    for h in range(1, len(want)):
        tg, tw = got[h].read(), want[h].read()
        assert sorted(tg.dtype.names) == sorted(tw.dtype.names)  # test_helpers.py:77-81: by name
        for name in tw.dtype.names:
            if not np.array_equal(tg[name], tw[name]):
                assert np.allclose(tg[name], tw[name], rtol=1e-5, atol=1e-8), (h, name)


COMMON = (" --rp-max +60.0 --rt-max +60.0 --nt 15 --nproc 1 --in-attributes " + DATA +
          "/test_delta/delta_attributes.fits.gz --in-dir " + DATA + "/test_delta/Delta_LYA/")


@pytest.mark.parametrize("script,double,flags,golden", [
    ("picca_cf", "cf", " --rp-min +0.0 --np 15 --remove-same-half-plate-close-pairs", "cf"),
    ("picca_dmat", "cf", " --rp-min +0.0 --np 15 --rej 0.99 --remove-same-half-plate-close-pairs"
     " --no-redshift-evolution", "dmat"),
    ("picca_xcf", "xcf", " --rp-min -60.0 --np 30 --z-evol-obj 1. --drq " + DATA +
     "/test_delta/cat.fits", "xcf"),
    ("picca_xdmat", "xcf", " --rp-min -60.0 --np 30 --rej 0.99 --z-evol-obj 1."
     " --no-redshift-evolution --drq " + DATA + "/test_delta/cat.fits", "xdmat"),
])
def test_unmodified_script_on_oracle_matches_golden_fits(tmp_path, script, double, flags, golden):
    _script_on_oracle(tmp_path, script, double, COMMON + flags, golden)


ANGL = (" --nproc 1 --in-attributes " + DATA + "/test_delta/delta_attributes.fits.gz --in-dir " +
        DATA + "/test_delta/Delta_LYA/")


@pytest.mark.parametrize("script,double,flags,golden", [
    # test_3_cor.py:206-227: wavelength-ratio x angle binning, cosmo=None deltas
    ("picca_cf_angl", "cf", "", "cf_angl"),
    # test_3_cor.py:611-634
    ("picca_xcf_angl", "xcf", " --z-evol-obj 1. --drq " + DATA + "/test_delta/cat.fits",
     "xcf_angl"),
])
def test_unmodified_angular_script_on_oracle_matches_golden_fits(tmp_path, script, double, flags,
                                                                 golden):
    """``ang_correlation`` branch (cf.py:186-208, :350-354; xcf.py:161-182, :293-295) against the
    reference's own golden files cf_angl.fits.gz / xcf_angl.fits.gz."""
    _script_on_oracle(tmp_path, script, double, ANGL + flags, golden)


def _script_on_oracle(tmp_path, script, double, flags, golden):
    import importlib
    from tests.refharness import load
    load.reference_modules()
    mod = importlib.import_module("picca.bin." + script)
    oracle_mod = importlib.reload(importlib.import_module("oracle." + double))
    saved = getattr(mod, double)
    setattr(mod, double, oracle_mod)  # the script now drives the oracle double of picca.cf/xcf
    out = str(tmp_path / (golden + ".fits.gz"))
    try:
        mod.main((flags + " --out " + out).split())
    finally:
        setattr(mod, double, saved)
    _compare_fits(out, DATA + "/test_cor/" + golden + ".fits.gz")


@pytest.mark.parametrize("name", ["small", "ragged", "per_r_par", "xcf_like", "empty_bin"])
def test_export_covariance_equals_live_reference(name):
    """oracle/export.py against the live utils.compute_cov / utils.smooth_cov
    (py/picca/utils.py:100-128, :153-249) on the seeded cases: same NumPy calls in the same
    order, so bit for bit."""
    from oracle import export as oexp
    from tests.golden import cases_export
    from tests.refharness import load
    _, _, _, _, utils = load.reference_modules()
    utils.userprint = lambda *a, **k: None
    cfg = cases_export.CASES[name]
    xi, we, rp, rt = cases_export.inputs(cfg)
    want = utils.compute_cov(xi, we)
    got = oexp.compute_cov(xi, we)
    assert np.array_equal(got, want)
    kw = dict(delta_r_trans=cfg["delta_r_trans"], delta_r_par=cfg["delta_r_par"],
              per_r_par=cfg.get("per_r_par", False))
    want_s = utils.smooth_cov(xi, we, rp, rt, covariance=want.copy(), **kw)
    got_s = oexp.smooth_cov(xi, we, rp, rt, covariance=want.copy(), **kw)
    assert np.array_equal(got_s, want_s)
    with np.errstate(all="ignore"):
        want_b = utils.compute_cov_boot(xi, we, nboots=40, seed=7)
        got_b = oexp.compute_cov_boot(xi, we, nboots=40, seed=7)
    assert np.array_equal(got_b, want_b, equal_nan=True)


@pytest.mark.parametrize("which", ["bundled", "bundled_image", "bundled_image_rebin3", "image",
                                   "imageblind", "sdss", "desi", "blind", "lin_rebin2"])
def test_read_deltas_equals_live_reference(which, tmp_path):
    """oracle/io.py against the live io.read_deltas (py/picca/io.py:383-512), bit for bit: the
    reference's bundled Delta_LYA directory (936 forests) and the generated cases."""
    from oracle import io as oio
    from tests.golden import cases_io
    from tests.refharness import load
    _, _, io, constants, _ = load.reference_modules()
    io.userprint = lambda *a, **k: None
    cosmo = constants.Cosmo(Om=0.315, Or=0., Ok=0., wl=-1., blinding="none", verbose=False)
    if which == "bundled":
        in_dir, attr = DATA + "/test_delta/Delta_LYA/", DATA + "/test_delta/delta_attributes.fits.gz"
    elif which.startswith("bundled_image"):
        in_dir = DATA + "/test_delta/Delta_LYA_image/"
        attr = DATA + "/test_delta/delta_attributes.fits.gz"
    elif which in cases_io.IMAGE_CASES:
        in_dir, attr = cases_io.write_image_case(str(tmp_path), which)
    else:
        in_dir, attr = cases_io.write_case(str(tmp_path), which.split("_")[0])
    kw = dict(cases_io.READ_KW)
    if "rebin" in which:  # io.py:362-378, data.py:657-686 (test_3_cor.py:288-316 uses 3)
        kw["rebin_factor"] = int(which[-1])
    want = io.read_deltas(in_dir, cosmo=cosmo, nproc=1, delta_attributes=attr, **kw)
    tables = (cosmo.get_r_comov.x, cosmo.get_r_comov.y, cosmo.get_dist_m.y)
    got = oio.read_deltas(in_dir.rstrip("/"), tables=tables, delta_attributes=attr, **kw)
    assert got[1:] == want[1:]
    a, b = cases_io.flatten(got[0]), cases_io.flatten(want[0])
    for k in a:
        assert np.array_equal(a[k], b[k]), k


@pytest.mark.parametrize("type_corr,x_corr", [("DD", False), ("xDD", True), ("DR", True)])
def test_co_bit_exact_on_bundled_catalogue(fixture_data, type_corr, x_corr):
    """oracle/co.py against the live picca.co (co.py:35-202) on the reference's bundled quasar
    catalogue (cat.fits, 1000 objects), the binning of test_3_cor.py:1084-1139."""
    import importlib
    from oracle import co as oco
    from tests.refharness import load
    _, _, _, _, utils = load.reference_modules()
    import picca.co
    co = importlib.reload(picca.co)
    co.userprint = lambda *a, **k: None
    cosmo, objs, z_min2 = fixture_data[3], fixture_data[4], fixture_data[5]
    cfg = dict(r_par_min=-80. if type_corr == "xDD" else 0., r_par_max=80., r_trans_max=80.,
               num_bins_r_par=40 if type_corr == "xDD" else 20, num_bins_r_trans=20, nside=16,
               type_corr=type_corr, x_correlation=x_corr, z_cut_min=0., z_cut_max=10.,
               objs=objs, objs2=objs if x_corr else None,
               ang_max=utils.compute_ang_max(cosmo, 80., z_min2, z_min2),
               num_data=sum(len(v) for v in objs.values()))
    total = 0
    for hp in sorted(objs)[:6]:
        res = []
        for mod in (co, oco):
            for k, v in cfg.items():
                setattr(mod, k, v)
            mod.lock, mod.counter = load.DummyLock(), load.DummyCounter()
            mod.fill_neighs([hp])
            res.append(mod.compute_xi([hp]))
        for a, b in zip(*res):
            assert np.array_equal(a, b)
        total += int(res[0][4].sum())
    assert total > 100


# ---- the rows of SURVEY 8f: the reference's unmodified scripts on the oracle doubles against the
# reference's own golden FITS (exact command lines of test_3_cor.py)
def test_metal_dmat_script_on_oracle_matches_golden_fits(tmp_path):
    """picca_metal_dmat.py (test_3_cor.py:351-381) driving oracle.cf.compute_metal_dmat."""
    import importlib
    from tests.refharness import load
    load.reference_modules()
    mod = importlib.import_module("picca.bin.picca_metal_dmat")
    ocf = importlib.reload(importlib.import_module("oracle.cf"))
    saved = mod.cf
    mod.cf = ocf
    out = str(tmp_path / "metal_dmat.fits.gz")
    try:
        mod.main((COMMON + " --rp-min +0.0 --np 15 --rej 0.99 --abs-igm SiIII(1207) --out "
                  + out).split())
    finally:
        mod.cf = saved
    _compare_fits(out, DATA + "/test_cor/metal_dmat.fits.gz")


def test_metal_xdmat_script_on_oracle_matches_golden_fits(tmp_path):
    """picca_metal_xdmat.py (test_3_cor.py:700-730) driving oracle.xcf.compute_metal_dmat."""
    import importlib
    from tests.refharness import load
    load.reference_modules()
    mod = importlib.import_module("picca.bin.picca_metal_xdmat")
    oxcf = importlib.reload(importlib.import_module("oracle.xcf"))
    saved = mod.xcf
    mod.xcf = oxcf
    out = str(tmp_path / "metal_xdmat.fits.gz")
    try:
        mod.main((COMMON + " --rp-min -60.0 --np 30 --rej 0.99 --z-evol-obj 1. --abs-igm SiIII(1207)"
                  " --drq " + DATA + "/test_delta/cat.fits --out " + out).split())
    finally:
        mod.xcf = saved
    _compare_fits(out, DATA + "/test_cor/metal_xdmat.fits.gz")


def test_co_script_on_oracle_matches_golden_fits(tmp_path):
    """picca_co.py --type-corr DD (test_3_cor.py:1084-1095) driving oracle.co."""
    import importlib
    from tests.refharness import load
    load.reference_modules()
    mod = importlib.import_module("picca.bin.picca_co")
    oco = importlib.reload(importlib.import_module("oracle.co"))
    saved = mod.co
    mod.co = oco
    out = str(tmp_path / "co_DD.fits.gz")
    try:
        mod.main(("--drq " + DATA + "/test_delta/cat.fits --out " + out + " --rp-min 0. --rp-max +60.0"
                  " --rt-max +60.0 --np 15 --nt 15 --nproc 1 --type-corr DD").split())
    finally:
        mod.co = saved
    _compare_fits(out, DATA + "/test_cor/co_DD.fits.gz")


def test_export_script_on_oracle_matches_golden_fits(tmp_path):
    """picca_export.py (test_3_cor.py:443-456) with the oracle's compute_cov / smooth_cov."""
    import importlib
    from oracle import export as oexp
    from tests.refharness import load
    load.reference_modules()
    mod = importlib.import_module("picca.bin.picca_export")
    saved = mod.compute_cov, mod.smooth_cov
    mod.compute_cov, mod.smooth_cov = oexp.compute_cov, oexp.smooth_cov
    out = str(tmp_path / "exported_cf.fits.gz")
    try:
        mod.main(("--data " + DATA + "/test_cor/cf.fits.gz --dmat " + DATA +
                  "/test_cor/dmat.fits.gz --out " + out).split())
    finally:
        mod.compute_cov, mod.smooth_cov = saved
    _compare_fits(out, DATA + "/test_cor/exported_cf.fits.gz")


@pytest.mark.parametrize("flags,golden", [("", "cf_image"), (" --rebin-factor 3", "cf_image_rebinned")])
def test_cf_script_on_oracle_loader_matches_golden_fits(tmp_path, flags, golden):
    """picca_cf.py on the ImageHDU deltas (test_3_cor.py:259-316) with BOTH the oracle loader
    (oracle.io.read_deltas in place of io.read_deltas) and the oracle cf."""
    import importlib
    from oracle import io as oio
    from tests.refharness import load
    load.reference_modules()
    mod = importlib.import_module("picca.bin.picca_cf")
    ocf = importlib.reload(importlib.import_module("oracle.cf"))

    def read_deltas(in_dir, nside, lambda_abs, alpha, z_ref, cosmo, max_num_spec=None,
                    no_project=False, nproc=None, rebin_factor=None, z_min_qso=0, z_max_qso=10,
                    delta_attributes=None):
        tables = (cosmo.get_r_comov.x, cosmo.get_r_comov.y, cosmo.get_dist_m.y)
        return oio.read_deltas(in_dir.rstrip("/"), nside, lambda_abs, alpha, z_ref, tables,
                               max_num_spec=max_num_spec, no_project=no_project,
                               z_min_qso=z_min_qso, z_max_qso=z_max_qso,
                               delta_attributes=delta_attributes, rebin_factor=rebin_factor)

    saved_cf, saved_rd = mod.cf, mod.io.read_deltas
    mod.cf, mod.io.read_deltas = ocf, read_deltas
    out = str(tmp_path / (golden + ".fits.gz"))
    try:
        mod.main((" --rp-max +60.0 --rt-max +60.0 --nt 15 --nproc 1 --rp-min +0.0 --np 15"
                  " --in-attributes " + DATA + "/test_delta/delta_attributes.fits.gz --in-dir " +
                  DATA + "/test_delta/Delta_LYA_image/" + flags + " --out " + out).split())
    finally:
        mod.cf, mod.io.read_deltas = saved_cf, saved_rd
    _compare_fits(out, DATA + "/test_cor/" + golden + ".fits.gz")


def test_wick_t123_on_bundled_fixtures(fixture_data):
    """oracle compute_wick_terms (max_diagram 3) against the live cf.compute_wick_terms /
    compute_wickT123_pairs (cf.py:1326-1626) on a few HEALPix pixels of the bundled forests, with
    analytic stand-ins for the 1-D products the script interpolates (picca_wick.py:393-416).
    SURVEY 8f rank 4; the CUDA counterpart is pb2_wick.cu (tests/test_wick_gpu.py)."""
    from tests.refharness import load
    from oracle import cf as ocf
    cf, _, _, _, utils = load.reference_modules()
    cf.userprint = lambda *a, **k: None
    hps = sorted(fixture_data[0])[:3]
    over = dict(r_par_max=20., r_trans_max=20., num_bins_r_par=5, num_bins_r_trans=5, reject=0.9,
                max_diagram=3,
                get_variance_1d={"D1": lambda ll: 0.05 + 0.01 * (ll - 3.55)},
                xi_1d={"D1": lambda dll: np.exp(-dll / 2e-3)})
    results = []
    for mod in (cf, ocf):
        _setup(mod, fixture_data, utils, load, **over)
        for k, v in over.items():
            setattr(mod, k, v)
        for hp in fixture_data[0]:
            for d in fixture_data[0][hp]:
                d.fname = "D1"  # picca_wick.py:371
        mod.fill_neighs(hps)
        np.random.seed(hps[0])
        results.append(mod.compute_wick_terms(hps))
    want, got = results
    assert (want[2], want[3]) == (got[2], got[3]) and want[3] > 0
    assert want[1].sum() > 0 and np.array_equal(want[1], got[1])
    assert np.array_equal(want[0], got[0])
    # the evolution factors are powers taken inside Numba (libm) there and by NumPy here: last ulp
    for k in (4, 5, 6):
        np.testing.assert_allclose(got[k], want[k], rtol=1e-13, atol=1e-13 * np.abs(want[k]).max())
        assert np.array_equal(got[k] != 0, want[k] != 0)
    assert np.abs(want[5]).sum() > 0 and np.abs(want[6]).sum() > 0


def test_xwick_t1234_on_bundled_fixtures(fixture_data):
    """oracle xcf.compute_wick_terms against the live xcf.compute_wick_terms /
    compute_wickT1234_pairs (xcf.py:838-1351) on the bundled forests x quasars, with the kind of
    interpolators picca_xwick.py builds (:394-409: interp1d, nearest, extrapolating)."""
    from scipy.interpolate import interp1d
    from tests.refharness import load
    from oracle import xcf as oxcf
    _, xcf, _, _, utils = load.reference_modules()
    xcf.userprint = lambda *a, **k: None
    hps = sorted(fixture_data[0])
    ll = 3.55 + 3e-4 * np.arange(400)
    over = dict(r_par_min=-60., num_bins_r_par=30, num_model_bins_r_par=30, alpha_obj=1.,
                reject=0.5, max_diagram=4, xi_wick=None,
                get_variance_1d={"D1": interp1d(ll, 0.05 + 0.1 * (ll - 3.55), kind="nearest",
                                                fill_value="extrapolate")},
                xi_1d={"D1": interp1d(ll - ll[0], np.exp(-(ll - ll[0]) / 2e-3), kind="nearest",
                                      fill_value="extrapolate")})
    results = []
    for mod in (xcf, oxcf):
        _setup(mod, fixture_data, utils, load, cross_obj=True, **over)
        for k, v in over.items():
            setattr(mod, k, v)
        for hp in fixture_data[0]:
            for d in fixture_data[0][hp]:
                d.fname = "D1"  # picca_xwick.py:336
        mod.fill_neighs(hps)
        np.random.seed(hps[0])
        results.append(mod.compute_wick_terms(hps))
    want, got = results
    assert (want[2], want[3]) == (got[2], got[3]) and want[3] > 100
    assert want[1].sum() > 1000 and np.array_equal(want[1], got[1])
    assert np.array_equal(want[0], got[0])
    for k in (4, 5, 6, 7):
        np.testing.assert_allclose(got[k], want[k], rtol=1e-13, atol=1e-13 * np.abs(want[k]).max())
        assert np.array_equal(got[k] != 0, want[k] != 0)
        assert np.abs(want[k]).sum() > 0


@pytest.mark.parametrize("script,double,flags,golden", [
    # test_3_cor.py:413-445 / :759-793 (default --max-diagram: 3 for the auto-, 4 for the
    # cross-correlation)
    ("picca_wick", "cf", " --rp-min +0.0 --np 15 --rej 0.99 --cf1d " + DATA +
     "/test_cor/cf1d.fits.gz", "wick"),
    ("picca_xwick", "xcf", " --rp-min -60.0 --np 30 --rej 0.99 --z-evol-obj 1. --cf1d " + DATA +
     "/test_cor/cf1d.fits.gz --drq " + DATA + "/test_delta/cat.fits", "xwick"),
])
def test_unmodified_wick_script_on_oracle_matches_golden_fits(tmp_path, script, double, flags,
                                                              golden):
    _script_on_oracle(tmp_path, script, double, COMMON + flags, golden)
