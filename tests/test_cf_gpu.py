"""GPU parity: picca_b200.cf (CUDA, through the C ABI) against the oracle double of picca.cf."""
import numpy as np
import pytest

from tests import helpers

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sample():
    from picca_b200 import synth
    data, num, z_min, z_max, cosmo = helpers.small_sample()
    ang_max = synth.compute_ang_max(cosmo, 60., z_min)
    return data, num, ang_max


@pytest.mark.parametrize("variant", [1, 2, 3, 0])
@pytest.mark.parametrize("over", [
    dict(),
    dict(remove_same_half_plate_close_pairs=True),
    dict(z_min_pairs=2.0, z_max_pairs=2.6),
    dict(zerr_cut_deg=0.5, zerr_cut_kms=40000.),
    dict(rmu_binning=True, r_par_min=0., r_par_max=1.),
    dict(num_bins_r_par=50, num_bins_r_trans=50, r_par_max=200., r_trans_max=200.),
])
def test_compute_xi_matches_oracle(sample, variant, over):
    from oracle import cf as ocf
    from picca_b200 import cf, synth
    data, num, ang_max = sample
    if over.get("r_trans_max", 60.) != 60.:
        ang_max = synth.compute_ang_max(synth.FlatLCDM(), over["r_trans_max"], 1.7)
    helpers.configure(ocf, data, num, ang_max, **over)
    helpers.configure(cf, data, num, ang_max, **over)
    cf._XI_VARIANT = variant
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
    assert total > 0
    cf._XI_VARIANT = 0


def test_forest_pair_functions_accumulate_in_place(sample):
    """compute_xi_forest_pairs_fast / compute_dmat_forest_pairs_fast keep the reference's
    positional signatures and in-place semantics (cf.py:251-272, :521-544)."""
    from oracle import _host, _kernels
    from oracle import cf as ocf
    from picca_b200 import cf
    data, num, ang_max = sample
    over = dict(reject=0., redshift_evolution_in_distortion_matrix=True)
    helpers.configure(ocf, data, num, ang_max, **over)
    helpers.configure(cf, data, num, ang_max, **over)
    hp = sorted(data)[0]
    ocf.fill_neighs([hp])
    d1 = next(d for d in data[hp] if len(d.neighbours) > 0)
    d2 = d1.neighbours[0]
    ang = float(_host.angle_between_one(d1, d2))
    p = _kernels.params_from_module(ocf)
    nb = 15 * 15
    want = [np.zeros(nb) for _ in range(5)] + [np.zeros(nb, dtype=np.int64)]
    _kernels.xi_auto_pair(p, d1, d2, ang, 0, want)
    got = [np.ones(nb) for _ in range(5)] + [np.ones(nb, dtype=np.int64)]
    cf.compute_xi_forest_pairs_fast(d1.z, d1.r_comov, d1.dist_m, d1.weights, d1.delta, d1.z_qso,
                                    d2.z, d2.r_comov, d2.dist_m, d2.weights, d2.delta, d2.z_qso,
                                    ang, False, *got)
    helpers.assert_xi_close([g - 1 for g in got], want, tag="pair fn")
    wd, dm = np.zeros(nb), np.zeros(nb * nb)
    eff = [np.zeros(nb) for _ in range(4)]
    _kernels.dmat_auto_pair(p, d1, d2, ang, 0, wd, dm, *eff)
    wd2, dm2 = np.zeros(nb), np.zeros(nb * nb)
    eff2 = [np.zeros(nb) for _ in range(4)]
    cf.compute_dmat_forest_pairs_fast(d1.log_lambda, d2.log_lambda, d1.r_comov, d2.r_comov,
                                      d1.dist_m, d2.dist_m, d1.z, d2.z, d1.weights, d2.weights,
                                      d1.z_qso, d2.z_qso, ang, wd2, dm2, *eff2, False, 1, 1)
    for a, b in zip([wd, dm] + eff, [wd2, dm2] + eff2):
        np.testing.assert_allclose(b, a, rtol=1e-9, atol=1e-12 * np.abs(a).max())
    for d in data[hp]:
        d.neighbours = None
