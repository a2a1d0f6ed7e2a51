"""CPU: host-side logic of the product that needs no GPU -- the delta loader (SoA/CSR packing,
bounding caps, sortedness), the parameter snapshot, the RNG stream contract, the generator."""
import numpy as np
import pytest

from picca_b200 import catalog, synth
from picca_b200.params import params_from_module
from tests import helpers


@pytest.fixture(scope="module")
def sample():
    return helpers.small_sample(n=200, seed=3, max_pix=60)


def test_pack_order_and_csr(sample):
    data = sample[0]
    cat = catalog.pack(data)
    flat = [d for hp in sorted(data) for d in data[hp]]
    assert cat.n_los == len(flat) == 200
    off = cat.arrays["offset"]
    assert off[0] == 0 and off[-1] == cat.n_pix == sum(len(d.weights) for d in flat)
    for k in (0, 17, 199):
        a, b = off[k], off[k + 1]
        assert np.array_equal(cat.arrays["r_comov"][a:b], flat[k].r_comov)
        assert np.array_equal(cat.arrays["delta_w"][a:b],
                              np.where(flat[k].weights != 0, flat[k].delta * flat[k].weights, 0.))
        assert cat.arrays["thingid"][k] == flat[k].thingid
    hp_first = cat.arrays["hp_first"]
    assert list(np.diff(hp_first)) == [len(data[hp]) for hp in sorted(data)]
    assert cat.sorted == 1 and cat.max_pix == 60


def test_caps_contain_members(sample):
    cat = catalog.pack(sample[0])
    A = cat.arrays
    for k in range(len(cat.healpixs)):
        a, b = A["hp_first"][k], A["hp_first"][k + 1]
        dots = (A["x_cart"][a:b] * A["cap_x"][k] + A["y_cart"][a:b] * A["cap_y"][k] +
                A["z_cart"][a:b] * A["cap_z"][k])
        assert np.all(np.arccos(np.clip(dots, -1, 1)) <= A["cap_rad"][k])


def test_unsorted_forest_is_flagged(sample):
    data = {hp: list(v) for hp, v in sample[0].items()}
    hp = sorted(data)[0]
    d = data[hp][0]
    import copy
    d2 = copy.copy(d)
    d2.r_comov = d.r_comov[::-1].copy()
    data[hp][0] = d2
    assert catalog.pack(data).sorted == 0


def test_zero_weight_nan_delta_does_not_leak(sample):
    import copy
    data = {hp: list(v) for hp, v in sample[0].items()}
    hp = sorted(data)[0]
    d = copy.copy(data[hp][0])
    d.weights, d.delta = d.weights.copy(), d.delta.copy()
    d.weights[3], d.delta[3] = 0., np.nan
    data[hp][0] = d
    assert np.all(np.isfinite(catalog.pack(data).arrays["delta_w"]))


def test_params_snapshot_reads_globals_at_call_time():
    class M:
        pass
    m = M()
    helpers.configure(m, {}, 0, 0.01)
    p = params_from_module(m)
    assert (p.num_bins_r_par, p.has_z_min_pairs, p.has_zerr_cut) == (15, 0, 0)
    m.z_min_pairs, m.zerr_cut_deg, m.zerr_cut_kms, m.num_bins_r_par = 2.1, 0.5, 400., 50
    p = params_from_module(m)
    assert (p.num_bins_r_par, p.has_z_min_pairs, p.z_min_pairs, p.has_zerr_cut) == (50, 1, 2.1, 1)
    m.r_par_max = None
    with pytest.raises(RuntimeError):
        params_from_module(m)


def test_rej_draw_stream_equals_per_forest_draws():
    """cf.py:444 draws rand(len(neighbours)) per forest; one rand(total) is the same stream."""
    lens = [3, 0, 11, 1, 0, 40]
    np.random.seed(1234)
    a = np.concatenate([np.random.rand(n) for n in lens])
    np.random.seed(1234)
    b = np.random.rand(sum(lens))
    assert np.array_equal(a, b)


def test_generator_is_deterministic():
    a = helpers.small_sample(n=50, seed=9, max_pix=40)[0]
    b = helpers.small_sample(n=50, seed=9, max_pix=40)[0]
    assert sorted(a) == sorted(b)
    for hp in a:
        for x, y in zip(a[hp], b[hp]):
            assert np.array_equal(x.delta, y.delta) and np.array_equal(x.weights, y.weights)


def test_ang2pix_ring_known_values():
    # pixel centres map back to their own index (nside 4, all 192 pixels), via the test shim
    from tests.refharness import shims
    nside = 8
    vec = shims.pix2vec(nside, np.arange(12 * nside * nside))
    theta = np.arccos(vec[:, 2])
    phi = np.arctan2(vec[:, 1], vec[:, 0])
    assert np.array_equal(synth.ang2pix_ring(nside, theta, phi), np.arange(12 * nside * nside))


def test_diag_copies_layout(sample):
    """The packed copies the diagonal-lane xi kernel reads (include/picca_b200.h, pb2_catalog):
    zero-weight pixels dropped, 48-byte records, natural order + copy interleaved by LANES, dummy
    pixels (weight 0, distance 1e299 / 1e300) everywhere else."""
    import copy
    data = {hp: list(v) for hp, v in sample[0].items()}
    hp0 = sorted(data)[0]
    d = copy.copy(data[hp0][0])
    d.weights = d.weights.copy()
    d.weights[[2, 5]] = 0.
    data[hp0][0] = d
    cat = catalog.pack(data)
    lanes, pad, row_pad, chunk = catalog.diag_layout()
    assert cat.dg_lanes == lanes and cat.dg_ok == 1
    A = cat.arrays
    flat = [x for hp in sorted(data) for x in data[hp]]
    dg_rec, il_rec = catalog.diag_records_host(cat)  # the specification of pb2_pack_diag
    rec = dg_rec.reshape(-1, 6)
    il = il_rec.reshape(lanes, cat.il_total, 6)
    assert A["dg_count"][0] == int((d.weights != 0).sum()) < len(d.weights) - 1
    assert cat.dg_max_pix == int(A["dg_count"].max())
    seen = np.zeros(il.shape[:2], dtype=bool)
    for f in (0, 1, 57, 199):
        x = flat[f]
        keep = x.weights != 0
        n = int(keep.sum())
        assert A["dg_count"][f] == n
        want = np.stack([x.r_comov[keep], x.dist_m[keep], x.weights[keep],
                         (x.delta * x.weights)[keep], 0.5 * x.z[keep], np.zeros(n)], axis=1)
        a = A["dg_offset"][f]
        assert np.array_equal(rec[a:a + n], want)
        # dummies after the forest: weight 0, distance 1e299
        assert np.all(rec[a + n:a + n + row_pad, 2:] == 0)
        assert np.all(rec[a + n:a + n + row_pad, :2] == catalog.DIAG_DUMMY_ROW)
        # interleaved copy: pixel j -> plane (j + pad) % lanes, record il_offset + (j + pad) // lanes
        jp = np.arange(n) + pad
        got = il[jp % lanes, A["il_offset"][f] + jp // lanes]
        assert np.array_equal(got, want)
        seen[jp % lanes, A["il_offset"][f] + jp // lanes] = True
        # pad dummies either side
        for j in (-pad, -1, n, n + pad - lanes):
            q = j + pad
            r = il[q % lanes, A["il_offset"][f] + q // lanes]
            assert r[2] == 0 and r[0] == catalog.DIAG_DUMMY_COL
    # a whole chunk of records may be read past the last forest in every plane
    last = A["il_offset"][-1] + (int(A["dg_count"][-1]) + 2 * pad + lanes - 1) // lanes
    assert cat.il_total >= last + chunk
    assert rec.shape[0] >= A["dg_offset"][-1] + A["dg_count"][-1] + row_pad + chunk


def test_diag_copies_flag_non_finite(sample):
    import copy
    data = {hp: list(v) for hp, v in sample[0].items()}
    hp0 = sorted(data)[0]
    d = copy.copy(data[hp0][0])
    d.z = d.z.copy()
    d.z[1] = np.inf
    data[hp0][0] = d
    assert catalog.pack(data).dg_ok == 0


def test_pair_list_subset_matches_numpy_selection():
    """PairList.subset (what np.array(neighbours)[w] selects, cf.py:444-447): CSR offsets and the
    per-pair arrays of the kept pairs, on CPU tensors."""
    import torch
    from picca_b200.engine import PairList
    counts = np.array([3, 0, 4, 1, 2])
    offset = np.concatenate([[0], np.cumsum(counts)])
    n = int(offset[-1])
    rng = np.random.default_rng(0)
    f1 = np.repeat(np.arange(len(counts)), counts).astype(np.int32)
    f2 = rng.integers(0, 50, n).astype(np.int32)
    ang = rng.random(n)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    pairs = PairList(None, t(np.arange(len(counts), dtype=np.int32)), t(offset), t(f1), t(f2),
                     t(ang), t(np.cos(ang / 2)), t(np.sin(ang / 2)))
    keep = rng.random(n) > 0.5
    sub = pairs.subset(keep)
    assert sub.n_f1 == pairs.n_f1 and sub.n_pairs == int(keep.sum())
    want_counts = np.array([keep[a:b].sum() for a, b in zip(offset[:-1], offset[1:])])
    assert np.array_equal(np.diff(sub.nb_offset.numpy()), want_counts)
    assert np.array_equal(sub.nb_f2.numpy(), f2[keep])
    assert np.array_equal(sub.nb_f1.numpy(), f1[keep])
    assert np.array_equal(sub.nb_ang.numpy(), ang[keep])


def test_smooth_cov_key_extents_cover_every_key():
    """export._smooth_extents sizes the device key table of pb2_cov_smooth: every key the
    reference's loops form (utils.py:203-211) must fall inside, for cf- and xcf-like binnings."""
    from picca_b200 import export
    from tests.golden import cases_export
    for name, cfg in cases_export.CASES.items():
        _, _, rp, rt = cases_export.inputs(cfg)
        for per_r_par in (False, True):
            n_dp, n_dt, rp_lo, n_rp = export._smooth_extents(rp, rt, cfg["delta_r_par"],
                                                             cfg["delta_r_trans"], per_r_par)
            i, j = np.triu_indices(len(rp), 1)
            k_dp = np.rint(np.abs(rp[j] - rp[i]) / cfg["delta_r_par"]).astype(int)
            k_dt = np.rint(np.abs(rt[i] - rt[j]) / cfg["delta_r_trans"]).astype(int)
            assert k_dp.max() < n_dp and k_dt.max() < n_dt, name
            if per_r_par:
                k_rp = np.trunc(rp[i] / cfg["delta_r_par"]).astype(int) - rp_lo
                assert k_rp.min() >= 0 and k_rp.max() < n_rp, name


def test_string_ids_share_one_mapping_across_catalogues():
    """Non-integer line-of-sight ids (combined re-observations) are mapped to integers through
    ONE process-wide table: the neighbour kernel compares ``thingid`` ACROSS catalogues (data vs
    data2, forests vs objects; cf.py:109-112, xcf.py:95-97), so equal ids must get equal
    integers in both and different ids different ones."""
    import copy
    from picca_b200 import catalog
    data, _, _, _, _ = helpers.small_sample(n=40, seed=3, max_pix=20)
    data2 = {hp: [copy.copy(d) for d in v] for hp, v in data.items()}
    for v in data.values():
        for d in v:
            d.thingid = "obj-%d" % d.thingid
    for k, v in enumerate(data2.values()):
        for d in reversed(v):      # other order, and one foreign id per pixel
            d.thingid = "obj-%d" % d.thingid
        v[0].thingid = "foreign-%d" % k
    c1, c2 = catalog.pack(data), catalog.pack(data2)
    assert c1.thingid_remapped and c2.thingid_remapped
    ids1 = {o.thingid: int(t) for o, t in zip(c1.objs, c1.arrays["thingid"])}
    ids2 = {o.thingid: int(t) for o, t in zip(c2.objs, c2.arrays["thingid"])}
    for name, t in ids2.items():
        if name in ids1:
            assert ids1[name] == t
        else:
            assert t not in ids1.values()
    assert len(set(ids1.values())) == len(ids1)


def test_cached_pack_sees_replaced_lists_and_invalidate():
    from picca_b200 import catalog
    data, _, _, _, _ = helpers.small_sample(n=30, seed=4, max_pix=20)
    a = catalog.cached_pack(data)
    assert catalog.cached_pack(data) is a
    hp = sorted(data)[0]
    data[hp] = list(reversed(data[hp]))          # same length, other list object
    b = catalog.cached_pack(data)
    assert b is not a and b.objs[0] is data[hp][0]
    data[hp][0].weights = data[hp][0].weights * 2.  # in-place style edit: invisible ...
    assert catalog.cached_pack(data) is b
    catalog.invalidate(data)                         # ... until the caller says so
    c = catalog.cached_pack(data)
    assert c is not b


def test_nearest_table_equals_scipy_interp1d():
    """The (bounds, values) table handed to the Wick kernels reproduces scipy's
    interp1d(kind="nearest", fill_value="extrapolate") -- picca_wick.py:412-417 -- including points
    exactly on a decision bound and beyond both ends."""
    import pytest
    from picca_b200 import _wick
    from tests.golden import cases
    _, xi = cases.wick_1d("D1")
    bounds, values = _wick.nearest_table(xi)
    x = np.abs(np.random.default_rng(0).normal(0., 0.05, 20000))
    x[:bounds.size] = np.asarray(xi.x_bds)
    x[-3:] = [0., 10., 1e-9]
    idx = np.searchsorted(bounds, x, side="left").clip(0, values.size - 1)
    assert np.array_equal(values[idx], xi(x))
    with pytest.raises(NotImplementedError):
        _wick.nearest_table(lambda v: v)


def test_forest_index_and_band_generation():
    """synth.make_forest_index / make_forest_band: a band is a function of (index, rows) only --
    two overlapping bands hold identical forests --, the deltas are projected (weighted mean and
    slope removed, data.py:622-655), the band packs from its registered SoA, and the index alone
    gives the RowIndex of the packed catalogue."""
    import numpy as np
    from picca_b200 import catalog, dist as pdist, synth
    ix = synth.make_forest_index(1200, seed=4, nside=32, ra_deg=(0., 20.), dec_deg=(0., 12.), max_pix=30)
    n_rows = len(ix.healpixs)
    assert int(ix.counts.sum()) == 1200 and ix.first[-1] == 1200
    whole = synth.make_forest_band(ix, 0, n_rows, threads=1)
    part = synth.make_forest_band(ix, n_rows // 3, n_rows // 3 + 5, rows_per_pass=2, threads=3)
    for hp in part:
        assert len(part[hp]) == len(whole[hp])
        for a, b in zip(part[hp], whole[hp]):
            assert a.thingid == b.thingid and a.ra == b.ra and a.z_qso == b.z_qso
            for name in ("delta", "weights", "log_lambda", "z", "r_comov", "dist_m"):
                assert np.array_equal(getattr(a, name), getattr(b, name))
    d = whole[ix.healpixs[n_rows // 2]][0]
    w = d.weights
    assert (w == 0).any() or True
    assert abs(np.sum(w * d.delta)) <= 1e-12 * np.sum(w)
    ll = d.log_lambda - np.average(d.log_lambda, weights=w)
    assert abs(np.sum(w * d.delta * ll)) <= 1e-12 * np.sum(w * np.abs(ll))
    assert np.allclose(np.diff(10**d.log_lambda), ix.dlambda)
    assert np.array_equal(d.z, 10**d.log_lambda / synth.LYA - 1.)
    host = catalog.pack(whole, defer_products=True)
    assert host.from_soa and host.n_los == 1200 and host.n_pix == int(ix.npix.sum())
    a, b = pdist.RowIndex(whole), pdist.RowIndex.from_arrays(ix.healpixs, ix.counts, ix.xyz, ix.npix)
    assert a.healpixs == b.healpixs and np.array_equal(a.first, b.first)
    assert np.array_equal(a.npix, b.npix)
    assert np.allclose(a.cap, b.cap, atol=1e-15) and np.allclose(a.cap_rad, b.cap_rad, atol=1e-12)
    zs = np.concatenate([o.z for v in whole.values() for o in v])
    assert abs(zs.min() - ix.z_min) < 1e-12 and abs(zs.max() - ix.z_max) < 1e-12
