"""Shared test helpers: configure a cf/xcf-like module, compare result tuples."""
import numpy as np

from picca_b200 import synth


class DummyLock:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


class DummyCounter:
    value = 0


CF_DEFAULTS = dict(
    num_bins_r_par=15, num_bins_r_trans=15, num_model_bins_r_par=15, num_model_bins_r_trans=15,
    r_par_max=60., r_par_min=0., r_trans_max=60., z_min_pairs=None, z_max_pairs=None,
    zerr_cut_deg=None, zerr_cut_kms=None, nside=16, z_ref=2.25, alpha=2.9, alpha2=2.9,
    x_correlation=False, rmu_binning=False, ang_correlation=False,
    remove_same_half_plate_close_pairs=False, redshift_evolution_in_distortion_matrix=True,
    reject=0.9, data2=None, num_data2=None)


def configure(mod, data, num_data, ang_max, **over):
    cfg = dict(CF_DEFAULTS)
    cfg.update(over)
    for key, val in cfg.items():
        if hasattr(mod, key) or key in CF_DEFAULTS:
            setattr(mod, key, val)
    mod.data = data
    mod.num_data = num_data
    mod.ang_max = ang_max
    mod.lock = DummyLock()
    mod.counter = DummyCounter()


def small_sample(n=300, seed=11, side_deg=6., max_pix=120, nside=16, **kw):
    data, num, z_min, z_max, cosmo = synth.make_forests(
        n, seed=seed, nside=nside, ra_deg=(10., 10. + side_deg), dec_deg=(5., 5. + side_deg),
        max_pix=max_pix, rest_range=kw.pop("rest_range", (1045., 1192.)), **kw)
    return data, num, z_min, z_max, cosmo


def assert_xi_close(got, want, rtol=1e-9, tag=""):
    """bit-exact num_pairs, fp64 sums within rtol (north_star tolerance 1e-9 relative)."""
    assert np.array_equal(np.asarray(got[5]), np.asarray(want[5])), \
        "%s num_pairs differ: %d vs %d" % (tag, np.sum(got[5]), np.sum(want[5]))
    names = ("weights", "xi", "r_par", "r_trans", "z")
    for k, name in enumerate(names):
        a, b = np.asarray(got[k]), np.asarray(want[k])
        scale = np.maximum(np.abs(b), 1e-300)
        # xi is a sum of signed terms: compare against the scale of the weighted terms
        if name == "xi":
            tol = rtol * np.maximum(np.abs(b), np.abs(b).max() * 1e-3 + 1e-300)
        else:
            tol = rtol * scale
        bad = np.abs(a - b) > tol
        assert not bad.any(), "%s %s: max rel err %.3e at %d bins" % (
            tag, name, (np.abs(a - b) / scale)[bad].max(), bad.sum())
