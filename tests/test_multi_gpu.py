"""GPU, >= 2 devices: the sharded paths (HEALPix rows LPT-partitioned over the ranks, blocks
gathered over NCCL; distortion matrix of one reference chunk with the kept forest pairs sharded
and ONE all-reduce) equal the single-GPU plugin calls -- num_pairs / NPALL / NPUSED exactly, sums
to 1e-11.  Skipped on a one-GPU box (run with ``gpurun --gpus 2``)."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_devices():
    import torch
    return torch.cuda.device_count()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2])
def test_sharded_xi_and_dmat_equal_single_gpu(world):
    if _n_devices() < world:
        pytest.skip("needs %d GPUs" % world)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
           "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "multi_gpu_worker.py")]
    res = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "multi-gpu ok" in res.stdout


def test_one_rank_shard_equals_plugin_call():
    """world = 1 through the same sharded entry points (no process group): the segmented --rej
    draw and the accumulate-into launches reproduce cf.compute_dmat."""
    import numpy as np
    from picca_b200 import catalog, cf, dist as pdist
    from picca_b200.engine import MODE_AUTO, get_engine
    from picca_b200.params import params_from_module
    from tests import helpers
    from tests.golden import cases
    eng = get_engine()
    cfg = dict(cases.DMAT_CASES["default"], reject=0.7)
    data, num, z_min, cosmo = cases.dmat_forests()
    ang_max = cases.ang_max_for(cosmo, cfg, z_min)
    helpers.configure(cf, data, num, ang_max, **cfg)
    host = catalog.cached_pack(data)
    dev = eng.device_catalog(host)
    params = params_from_module(cf)
    shard = pdist.Shard(eng, host, host, ang_max, 1, 0)
    hps = host.healpixs
    res, npall, npused = pdist.dmat_chunk_sharded(eng, dev, dev, params, shard, MODE_AUTO,
                                                  cf.reject, hps[0], segments=4)
    cf.fill_neighs(hps)
    np.random.seed(hps[0])
    one = cf.compute_dmat(hps)
    assert (npall, npused) == (one[6], one[7])
    for a, b in zip(res, one[:6]):
        a = a.cpu().numpy()
        assert np.abs(a - b).max() <= 1e-11 * max(np.abs(b).max(), 1e-300)
    full = pdist.xi_sharded(eng, dev, dev, params, shard, MODE_AUTO).cpu().numpy()
    cf.fill_neighs(hps)
    want = cf.compute_xi_batch(hps)
    assert np.array_equal(full[:, 5].view(np.int64), want[:, 5].view(np.int64))
    # the band-shard entry points with one band (= everything, no halo)
    band = pdist.BandShard(eng, data, ang_max, 1, 0)
    assert (band.b0, band.b1, band.h0, band.h1) == (0, len(hps), 0, len(hps))
    full_b = pdist.xi_banded(eng, band, params, MODE_AUTO).cpu().numpy()
    assert np.array_equal(full_b[:, 5].view(np.int64), want[:, 5].view(np.int64))
    res_b, npall_b, npused_b = pdist.dmat_chunk_banded(eng, band, params, MODE_AUTO, cf.reject,
                                                       hps[0], segments=3)
    assert (npall_b, npused_b) == (one[6], one[7])
    for a, b in zip(res_b, one[:6]):
        assert np.abs(a.cpu().numpy() - b).max() <= 1e-11 * max(np.abs(b).max(), 1e-300)


def test_band_source_shards_equal_the_whole_catalogue():
    """Band shards from a survey index + a band source (``synth.make_forest_index`` /
    ``make_forest_band``: no rank generates or holds more than its band + halo): three "ranks",
    run one after the other on this GPU, reproduce the rows of the whole catalogue."""
    import numpy as np
    from picca_b200 import cf, dist as pdist, synth
    from picca_b200.engine import MODE_AUTO, get_engine
    from picca_b200.params import params_from_module
    from tests import helpers
    eng = get_engine()
    ix = synth.make_forest_index(1800, seed=11, nside=32, ra_deg=(0., 24.), dec_deg=(0., 14.),
                                 max_pix=48)
    ang_max = synth.compute_ang_max(ix.cosmo, 200., ix.z_min)
    idx = pdist.RowIndex.from_arrays(ix.healpixs, ix.counts, ix.xyz, ix.npix)
    data = synth.make_forest_band(ix, 0, len(ix.healpixs))
    helpers.configure(cf, data, ix.n_forest, ang_max, num_bins_r_par=50, num_bins_r_trans=50,
                      r_par_max=200., r_trans_max=200., nside=32)
    hps = sorted(data)
    assert hps == ix.healpixs
    cf.fill_neighs(hps)
    want = cf.compute_xi_batch(hps)
    params = params_from_module(cf)
    world = 3
    seen = 0
    for rank in range(world):
        band = pdist.BandShard(eng, None, ang_max, world, rank, index=idx,
                               band_source=lambda h0, h1: synth.make_forest_band(ix, h0, h1))
        assert band.host.n_los == ix.first[band.h1] - ix.first[band.h0]
        got = pdist.xi_banded(eng, band, params, MODE_AUTO, gather=False).cpu().numpy()
        ref = want[band.b0:band.b1]
        assert np.array_equal(got[:, 5].view(np.int64), ref[:, 5].view(np.int64))
        for k in range(5):
            assert np.abs(got[:, k] - ref[:, k]).max() <= 1e-11 * max(np.abs(ref[:, k]).max(), 1e-300)
        seen += band.b1 - band.b0
    assert seen == len(hps)
    assert want[:, 5].view(np.int64).sum() > 0
