"""CPU: the plumbing of bench.py's CPU arms.  The sub-pixel tasks entered as extra keys of the
reference's ``cf.data`` produce, through the UNMODIFIED reference's fill_neighs + compute_xi
(Numba, fork pool -- the call pattern of picca_cf.py:449-463), exactly the rows the oracle port
computes for the same forests: num_pairs bit for bit, sums to 1e-9."""
import numpy as np
import pytest

from tests.refharness import shims

pytestmark = [pytest.mark.reference,
              pytest.mark.skipif(not shims.reference_available(), reason="needs the reference")]


def test_numba_reference_tasks_equal_oracle_port_tasks():
    import bench
    from picca_b200 import synth
    data, num, z_min, _, cosmo = synth.make_forests(1500, seed=5, nside=32, ra_deg=(0., 14.),
                                                    dec_deg=(0., 5.), max_pix=150)
    ang_max = synth.compute_ang_max(cosmo, 200., z_min)
    hps = sorted(data)
    tasks = bench.sample_tasks(data, hps, 6, 2, step=1)
    ref = bench.NumbaReference(data, num, ang_max)
    ref.warm_jit()
    rows_ref, dt = ref.run_tasks(tasks, 2)
    w = rows_ref[:, 0] > 0
    soa = bench.OracleSoA(data)
    cfg = bench.Cfg()
    bench.configure(cfg, data, num, ang_max)
    rows_port, _ = bench.port_run_tasks(soa, cfg, ang_max, tasks, 2)
    for k in (1, 2, 3, 4):
        rows_port[:, k][w] /= rows_port[:, 0][w]
    assert np.array_equal(rows_ref[:, 5].view(np.int64), rows_port[:, 5].view(np.int64))
    assert rows_ref[:, 5].view(np.int64).sum() > 10**6
    # the port adds the forests of a task on several threads: another summation order
    for k in range(5):
        np.testing.assert_allclose(rows_ref[:, k], rows_port[:, k], rtol=1e-9,
                                   atol=1e-12 * np.abs(rows_port[:, k]).max())
    assert set(ref.cf.data) == set(data)   # the task keys are gone again
