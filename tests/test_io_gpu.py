"""GPU: the delta loader (picca_b200.io.read_deltas -> pb2_fits_scan/pb2_fits_cards on the host,
pb2_delta_unpack / pb2_delta_prepare on the device, through the C ABI) against the live
reference's io.read_deltas outputs (tests/golden/golden_io.npz) and the oracle restatement.

Parity bar: identifiers, HEALPix ids, pixel counts, log_lambda bit-equal; with the host power
(PICCA_B200_HOST_POW=1) z, r_comov, dist_m, z_min, z_max bit-equal (the interpolation follows
scipy's interp1d operation by operation); with the device exp10 they agree within 4 ulp.
Evolved weights and projected deltas (pow, re-associated weighted sums): 1e-12 relative to the
scale of the forest's values."""
import os

import numpy as np
import pytest

from tests.golden import cases_io

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RUNS = {"fixture": None, "imagefixture": None, "image": {}, "imageblind": {}, "sdss": {}, "desi": {}, "blind": {},
        "sdss_noproject": dict(no_project=True), "sdss_max30": dict(max_num_spec=30),
        "sdss_zcut": dict(z_min_qso=2.4, z_max_qso=3.0),
        "lin_rebin2": dict(rebin_factor=2), "image_rebin3": dict(rebin_factor=3),
        "imagefixture_rebin3": dict(rebin_factor=3)}


class TableCosmo:
    """the reference Cosmo's tables (stored with the golden vectors)"""

    def __init__(self, gold):
        self._t = (gold["cosmo_z"], gold["cosmo_r_comov"], gold["cosmo_dist_m"])

    def table(self):
        return self._t


def inputs(tag, tmp_path):
    fx = os.path.join(GOLD, "fixtures")
    if tag == "fixture":
        return os.path.join(fx, "delta-272.fits.gz"), os.path.join(fx, "delta_attributes.fits.gz")
    if tag.startswith("imagefixture"):
        return (os.path.join(fx, "image-delta-50.fits.gz"),
                os.path.join(fx, "delta_attributes.fits.gz"))
    if tag.split("_")[0] in cases_io.IMAGE_CASES:
        return cases_io.write_image_case(str(tmp_path), tag.split("_")[0])
    return cases_io.write_case(str(tmp_path), tag.split("_")[0])


def ulp_distance(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.abs(a.view(np.int64) - b.view(np.int64))


def close_on_forest_scale(got, want, n_pix, rtol, atol=1e-300):
    edges = np.concatenate([[0], np.cumsum(n_pix)])
    for a, b in zip(edges[:-1], edges[1:]):
        if b > a:
            scale = np.abs(want[a:b]).max()
            assert np.all(np.abs(got[a:b] - want[a:b]) <= rtol * scale + atol)


@pytest.mark.parametrize("host_pow", [True, False])
@pytest.mark.parametrize("tag", sorted(RUNS))
def test_read_deltas_matches_reference_golden(tag, host_pow, tmp_path, monkeypatch):
    from picca_b200 import io
    io.userprint = lambda *a, **k: None
    monkeypatch.setenv("PICCA_B200_HOST_POW", "1" if host_pow else "0")
    gold = np.load(os.path.join(GOLD, "golden_io.npz"))
    in_dir, attr = inputs(tag, tmp_path)
    data, num, z_min, z_max = io.read_deltas(in_dir, cosmo=TableCosmo(gold), delta_attributes=attr,
                                             **dict(cases_io.READ_KW, **(RUNS[tag] or {})))
    flat = cases_io.flatten(data)
    g = lambda k: gold["%s_%s" % (tag, k)]
    assert num == int(g("summary")[0])
    for k in ("healpix", "los_id", "plate", "mjd", "fiberid", "n_pix", "order", "ra", "dec",
              "z_qso"):
        assert np.array_equal(flat[k], g(k)), k
    # files that store LAMBDA: log10 is taken by the loader (data.py:411-412)
    stores_lambda = (cases_io.CASES.get(tag.split("_")[0], {}).get("wave") == "LAMBDA" or
                     tag.startswith("image"))
    if host_pow:
        for k in ("log_lambda", "z", "r_comov", "dist_m"):
            assert np.array_equal(flat[k], g(k)), k
        assert [z_min, z_max] == list(g("summary")[1:])
    else:
        # rebinned forests: exp10 -> bin mid-points -> log10 all on the device
        rebinned = "rebin" in tag
        ll_ulp = 4 if rebinned else (2 if stores_lambda else 0)
        assert ulp_distance(flat["log_lambda"], g("log_lambda")).max() <= ll_ulp
        for k in ("z", "r_comov", "dist_m"):
            assert ulp_distance(flat[k], g(k)).max() <= (32 if rebinned else
                                                         16 if stores_lambda else 4), k
        np.testing.assert_allclose([z_min, z_max], g("summary")[1:], rtol=4e-15)
    close_on_forest_scale(flat["weights"], g("weights"), flat["n_pix"], 1e-12)
    # a projected delta is a difference of O(1) numbers (and a 3-pixel linear fit is a cancelling
    # sum): 1e-12 absolute on the scale of the unprojected deltas
    close_on_forest_scale(flat["delta"], g("delta"), flat["n_pix"], 1e-12, atol=1e-12)


def test_loader_feeds_the_pair_kernels_like_the_reference_loader(tmp_path, monkeypatch):
    """files -> picca_b200.io.read_deltas -> picca_b200.cf.compute_xi equals files -> oracle
    loader -> oracle cf: bit-exact num_pairs (host power), sums within 1e-9."""
    from oracle import cf as ocf, io as oio
    from picca_b200 import cf, io, synth
    from tests import helpers
    io.userprint = lambda *a, **k: None
    monkeypatch.setenv("PICCA_B200_HOST_POW", "1")
    gold = np.load(os.path.join(GOLD, "golden_io.npz"))
    cosmo = TableCosmo(gold)
    in_dir, attr = cases_io.write_case(str(tmp_path), "sdss")
    got = io.read_deltas(in_dir, cosmo=cosmo, delta_attributes=attr, **cases_io.READ_KW)
    want = oio.read_deltas(in_dir, tables=cosmo.table(), delta_attributes=attr, **cases_io.READ_KW)
    dist_min = np.interp(got[2], cosmo.table()[0], cosmo.table()[2])
    ang_max = float(2. * np.arcsin(60. / (2. * dist_min)))
    helpers.configure(ocf, want[0], want[1], ang_max, nside=cases_io.NSIDE)
    helpers.configure(cf, got[0], got[1], ang_max, nside=cases_io.NSIDE)
    total = 0
    for hp in sorted(got[0]):
        ocf.fill_neighs([hp])
        cf.fill_neighs([hp])
        a, b = cf.compute_xi([hp]), ocf.compute_xi([hp])
        helpers.assert_xi_close(a, b, tag="hp %d" % hp)
        total += int(np.sum(a[5]))
    assert total > 0


def test_error_behaviour(tmp_path):
    from picca_b200 import io
    io.userprint = lambda *a, **k: None
    gold = np.load(os.path.join(GOLD, "golden_io.npz"))
    cosmo = TableCosmo(gold)
    in_dir, attr = cases_io.write_case(str(tmp_path), "desi")
    with pytest.raises(ValueError):  # rebinning empties the single-pixel forests of this case
        io.read_deltas(in_dir, cosmo=cosmo, delta_attributes=attr, rebin_factor=2,
                       **cases_io.READ_KW)
    fx = os.path.join(GOLD, "fixtures")
    with pytest.raises(KeyError):  # no WAVE_SOLUTION card in the bundled BinTable file (io.py:368)
        io.read_deltas(os.path.join(fx, "delta-272.fits.gz"), cosmo=cosmo, rebin_factor=2,
                       delta_attributes=os.path.join(fx, "delta_attributes.fits.gz"),
                       **cases_io.READ_KW)
    with pytest.raises(AssertionError):  # io.py:489-490: nothing passes the quasar redshift cut
        io.read_deltas(in_dir, cosmo=cosmo, delta_attributes=attr, z_min_qso=8., z_max_qso=9.,
                       **cases_io.READ_KW)
    with pytest.raises(RuntimeError):  # data.py:628-633: projecting without a continuum order
        io.read_deltas(in_dir, cosmo=cosmo, delta_attributes=str(tmp_path / "missing.fits.gz"),
                       **cases_io.READ_KW)
    with pytest.raises(ValueError):  # redshifts beyond the cosmology table (interp1d bounds)
        io.read_deltas(in_dir, cosmo=cosmo, delta_attributes=attr,
                       **dict(cases_io.READ_KW, lambda_abs=200.))
    # cosmo=None: no distances (io.py:500)
    data, _, _, _ = io.read_deltas(in_dir, cosmo=None, delta_attributes=attr, **cases_io.READ_KW)
    d = next(iter(data.values()))[0]
    assert d.r_comov is None and d.z is not None
