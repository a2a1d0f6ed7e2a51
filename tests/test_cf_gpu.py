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


@pytest.mark.parametrize("variant", [1, 0])
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
