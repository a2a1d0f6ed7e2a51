"""GPU parity on the edge cases the reference's loops handle implicitly: forests with no usable
pixel, single-pixel forests, ragged lengths, forests longer than any tile, unsorted pixels
(falls back to the general kernel), duplicate sky positions (small-angle branch of
get_angle_between, data.py:158-161), empty neighbour lists, production binning for a
delta x delta cross-correlation, and the maximum bin counts of the xcf defaults."""
import copy

import numpy as np
import pytest

from tests import helpers

pytestmark = pytest.mark.gpu


def run_both(data, num, ang_max, **over):
    from oracle import cf as ocf
    from picca_b200 import cf
    helpers.configure(ocf, data, num, ang_max, **over)
    helpers.configure(cf, data, num, ang_max, **over)
    total = 0
    for hp in sorted(data):
        ocf.fill_neighs([hp])
        want_n = [[d2.thingid for d2 in d.neighbours] for d in data[hp]]
        want = ocf.compute_xi([hp])
        cf.fill_neighs([hp])
        got_n = [[d2.thingid for d2 in d.neighbours] for d in data[hp]]
        assert got_n == want_n
        got = cf.compute_xi([hp])
        helpers.assert_xi_close(got, want, tag="hp %d" % hp)
        total += int(want[5].sum())
    return total


@pytest.fixture()
def base():
    from picca_b200 import synth
    data, num, z_min, _, cosmo = helpers.small_sample(n=120, seed=77, max_pix=90, side_deg=4.)
    data = {hp: [copy.copy(d) for d in v] for hp, v in data.items()}
    return data, num, synth.compute_ang_max(cosmo, 60., z_min)


def test_zero_weight_and_tiny_forests(base):
    data, num, ang_max = base
    flat = [d for hp in sorted(data) for d in data[hp]]
    flat[0].weights = np.zeros_like(flat[0].weights)            # nothing usable
    for name in ("weights", "delta", "z", "r_comov", "dist_m", "log_lambda"):
        setattr(flat[1], name, getattr(flat[1], name)[:1].copy())  # one pixel
        setattr(flat[2], name, getattr(flat[2], name)[:2].copy())  # two pixels
    flat[3].weights = flat[3].weights.copy()
    flat[3].weights[::2] = 0.                                   # every other pixel masked
    assert run_both(data, num, ang_max) > 0


def test_ragged_and_long_forests():
    from picca_b200 import synth
    # forests up to ~1400 pixels: longer than one row tile / several diagonal blocks
    data, num, z_min, _, cosmo = synth.make_forests(
        40, seed=5, nside=16, ra_deg=(10., 12.), dec_deg=(5., 7.), rest_range=(1000., 1250.),
        dlambda=0.4)
    lens = [len(d.weights) for hp in data for d in data[hp]]
    assert max(lens) > 1100 and min(lens) < max(lens)
    ang_max = synth.compute_ang_max(cosmo, 60., z_min)
    assert run_both(data, num, ang_max) > 0
    assert run_both(data, num, ang_max, num_bins_r_par=50, num_bins_r_trans=50, r_par_max=200.,
                    r_trans_max=200.) > 0


def test_unsorted_forest_uses_general_kernel(base):
    data, num, ang_max = base
    flat = [d for hp in sorted(data) for d in data[hp]]
    d = flat[5]
    perm = np.random.default_rng(0).permutation(len(d.weights))
    for name in ("weights", "delta", "z", "r_comov", "dist_m", "log_lambda"):
        setattr(d, name, getattr(d, name)[perm].copy())
    from picca_b200 import catalog
    assert catalog.pack(data).sorted == 0
    assert run_both(data, num, ang_max) > 0


def test_duplicate_positions_small_angle_branch(base):
    data, num, ang_max = base
    hp = sorted(data)[0]
    a = data[hp][0]
    twin = copy.copy(a)
    twin.thingid = twin.los_id = 999001
    twin.ra = a.ra + 1e-7          # within 2 arcsec: sqrt(ddec^2 + (cos_dec dra)^2) branch
    twin.dec = a.dec - 2e-7
    twin.x_cart = np.cos(twin.ra) * np.cos(twin.dec)
    twin.y_cart = np.sin(twin.ra) * np.cos(twin.dec)
    twin.z_cart = np.sin(twin.dec)
    twin.cos_dec = np.cos(twin.dec)
    data[hp].append(twin)
    assert run_both(data, num + 1, ang_max) > 0


def test_isolated_forest_has_no_neighbours():
    from picca_b200 import cf, synth
    data, num, z_min, _, cosmo = helpers.small_sample(n=3, seed=1, max_pix=30, side_deg=30.)
    ang_max = synth.compute_ang_max(cosmo, 60., z_min)
    helpers.configure(cf, data, num, ang_max)
    for hp in sorted(data):
        cf.fill_neighs([hp])
        res = cf.compute_xi([hp])
        assert all(len(r) == 225 for r in res)
    # at least one of the three widely separated forests has an empty list and gives zeros
    assert any(int(cf.compute_xi_batch([hp])[0, 5].view(np.int64).sum()) == 0
               for hp in sorted(data) if not cf.fill_neighs([hp]))


def test_cross_correlation_production_binning(base):
    data, num, ang_max = base
    from picca_b200 import synth
    data2, num2, z_min2, _, cosmo = helpers.small_sample(n=90, seed=78, max_pix=80, side_deg=4.,
                                                        id_offset=7000)
    ang_max = synth.compute_ang_max(cosmo, 200., 1.96, z_min2)
    total = run_both(data, num, ang_max, data2=data2, num_data2=num2, x_correlation=True,
                     r_par_min=-200., r_par_max=200., r_trans_max=200., num_bins_r_par=100,
                     num_bins_r_trans=50)
    assert total > 10**6


def _grid_forest(los_id, ra, dec, first, n, rng):
    """A forest whose comoving distances sit on a 0.5 Mpc/h grid: with bins of 4 Mpc/h many pixel
    pairs fall EXACTLY on r_par bin edges (and on r_par = 0)."""
    from picca_b200.forest import Delta
    rc = 3000. + 0.5 * (first + np.arange(n, dtype=np.float64))
    d = Delta(los_id, ra, dec, 2.5, los_id, los_id, los_id, np.log10(3600. + np.arange(n)),
              rng.uniform(0.5, 2., n), rng.normal(0., 0.3, n), 1)
    d.weights[rng.random(n) < 0.1] = 0.
    d.z = 2. + 1e-3 * (first + np.arange(n, dtype=np.float64))
    d.r_comov = rc
    d.dist_m = rc.copy()
    return d


def test_pairs_exactly_on_bin_edges():
    """Bin-edge pairs: the diagonal-lane kernel must take the reference expression for them
    (fraction 0 or 65535 of its fixed-point bin value) and still count every pair once."""
    rng = np.random.default_rng(4)
    forests = []
    # nearly the same sky position (ang ~ 1e-9: cos(ang/2) rounds to 1 -> r_par exactly on the
    # grid, r_trans ~ 0), 1e-4 rad neighbours, and shifted grids so that forests overlap partially
    for k, (dra, ddec, first, n) in enumerate([(0., 0., 0, 150), (1e-9, 0., 40, 97), (1e-4, 0., 3, 200),
                                               (2e-4, 1e-4, 90, 130), (2e-9, 1.5e-4, 16, 64)]):
        forests.append(_grid_forest(500 + k, 0.3 + dra, 0.1 + ddec, first, n, rng))
    data = {7: forests}
    ang_max = 1e-3
    total = run_both(data, len(forests), ang_max, num_bins_r_par=15, num_bins_r_trans=15,
                     r_par_max=60., r_trans_max=60.)
    assert total > 10**4
    # production binning (4 Mpc/h bins, every eighth grid point is an edge)
    total = run_both(data, len(forests), ang_max, num_bins_r_par=50, num_bins_r_trans=50,
                     r_par_max=200., r_trans_max=200.)
    assert total > 10**4


def test_device_packed_copies_equal_the_numpy_specification():
    """pb2_pack_diag (records written in HBM from the SoA) against catalog.diag_records_host, bit
    for bit, on a sample with zero-weight pixels, ragged and tiny forests."""
    from picca_b200 import catalog
    from picca_b200.engine import get_engine
    data, num, z_min, _, cosmo = helpers.small_sample(n=260, seed=3, max_pix=150)
    rng = np.random.default_rng(1)
    flat = [d for hp in sorted(data) for d in data[hp]]
    for k in (0, 5, 77):
        flat[k].weights = flat[k].weights.copy()
        flat[k].weights[rng.random(len(flat[k].weights)) < 0.4] = 0.
    flat[9].weights = np.zeros_like(flat[9].weights)  # a forest that disappears entirely
    host = catalog.pack(data)
    eng = get_engine()
    dev = eng.device_catalog(host, cache=False)
    want_dg, want_il = catalog.diag_records_host(host)
    assert np.array_equal(dev.tensors["dg_rec"].cpu().numpy(), want_dg)
    assert np.array_equal(dev.tensors["il_rec"].cpu().numpy(), want_il)
    assert host.arrays["dg_count"][9] == 0


def test_string_ids_across_two_catalogues(base):
    """delta x delta cross-correlation whose thingids are strings: a forest must still be
    excluded from pairing with ITS OWN copy in the other catalogue (cf.py:109-112), which needs
    one id -> integer mapping shared by both packed catalogues."""
    data, num, ang_max = base
    data2 = {hp: [copy.copy(d) for d in v] for hp, v in data.items()}
    for cat in (data, data2):
        for v in cat.values():
            for d in v:
                d.thingid = "tid-%s" % d.thingid
    hp0 = sorted(data2)[0]
    data2[hp0] = list(reversed(data2[hp0]))   # other insertion order than `data`
    from oracle import cf as ocf
    from picca_b200 import cf
    over = dict(data2=data2, num_data2=num, x_correlation=True, r_par_min=-60., num_bins_r_par=30)
    total = 0
    for mod in (ocf, cf):
        helpers.configure(mod, data, num, ang_max, **over)
    for hp in sorted(data):
        ocf.fill_neighs([hp])
        want_n = [[d2.thingid for d2 in d.neighbours] for d in data[hp]]
        want = ocf.compute_xi([hp])
        cf.fill_neighs([hp])
        assert [[d2.thingid for d2 in d.neighbours] for d in data[hp]] == want_n
        assert all(d.thingid not in ids for d, ids in zip(data[hp], want_n))
        helpers.assert_xi_close(cf.compute_xi([hp]), want, tag="hp %d" % hp)
        total += int(want[5].sum())
    assert total > 0
    cf.data2 = ocf.data2 = None


@pytest.mark.parametrize("spoil", ["none", "unsorted", "nan_kept", "nan_zero_weight"])
def test_deferred_pack_equals_host_pack(base, spoil):
    """The product path leaves delta*weights, z*weights, the counts of non-zero weights and the
    finiteness / sortedness / reach flags to the device (pb2_derive_products, pb2_catalog_stats):
    same arrays, same layout, same flags as the NumPy packing."""
    from picca_b200 import catalog
    from picca_b200.engine import get_engine
    data, num, ang_max = base
    hp = sorted(data)[1]
    d = data[hp][0]
    if spoil == "unsorted":
        d.r_comov = d.r_comov[::-1].copy()
    elif spoil == "nan_kept":
        d.delta = d.delta.copy()
        d.delta[np.nonzero(d.weights)[0][0]] = np.nan
    elif spoil == "nan_zero_weight":
        d.weights = d.weights.copy()
        d.delta = d.delta.copy()
        d.weights[3] = 0.
        d.delta[3] = np.nan
    want = catalog.pack(data)
    got = catalog.pack(data, defer_products=True)
    assert got.meta_deferred and "delta_w" not in got.arrays
    dev = get_engine().device_catalog(got, cache=False)
    assert not got.meta_deferred
    for name in ("dg_offset", "dg_count", "il_offset"):
        assert np.array_equal(got.arrays[name], want.arrays[name]), name
    for attr in ("dg_total", "il_total", "dg_max_pix", "dg_ok", "sorted", "dg_lanes"):
        assert getattr(got, attr) == getattr(want, attr), attr
    assert got.dg_reach == want.dg_reach
    for name in ("delta_w", "z_w"):
        assert np.array_equal(dev.tensors[name].cpu().numpy(), want.arrays[name], equal_nan=True), name
    assert (want.sorted, want.dg_ok) == {"none": (1, 1), "unsorted": (0, 1), "nan_kept": (1, 0),
                                         "nan_zero_weight": (1, 1)}[spoil]
