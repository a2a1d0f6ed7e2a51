"""Worker of tests/test_multi_gpu.py, launched with torchrun (one process per GPU, NCCL):
sharded xi + sharded distortion matrix of a seeded sample; rank 0 compares them with the
single-process plugin calls (picca_b200.cf.compute_xi_batch / compute_dmat) on its own GPU."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from picca_b200 import catalog, cf, dist as pdist
    from picca_b200.engine import MODE_AUTO, get_engine
    from picca_b200.params import params_from_module
    from tests import helpers
    from tests.golden import cases

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ["LOCAL_RANK"])
    os.environ["PICCA_B200_DEVICE"] = str(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    eng = get_engine()
    cfg = dict(cases.DMAT_CASES["default"], reject=0.7)
    data, num, z_min, _, cosmo = helpers.small_sample(n=400, seed=41, max_pix=90, side_deg=8.)
    ang_max = cases.ang_max_for(cosmo, cfg, z_min)
    helpers.configure(cf, data, num, ang_max, **cfg)
    host = catalog.cached_pack(data)
    dev = eng.device_catalog(host)
    params = params_from_module(cf)
    shard = pdist.Shard(eng, host, host, ang_max, world, rank)
    assert sorted(np.concatenate(shard.parts).tolist()) == list(range(len(host.healpixs)))
    hps = host.healpixs

    full = pdist.xi_sharded(eng, dev, dev, params, shard, MODE_AUTO)
    res, npall, npused = pdist.dmat_chunk_sharded(eng, dev, dev, params, shard, MODE_AUTO,
                                                  cf.reject, hps[0], segments=5)
    # band shards: each rank packs and uploads only its band of rows + halo
    band = pdist.BandShard(eng, data, ang_max, world, rank)
    assert band.host.n_los <= host.n_los and band.host.from_soa
    full_b = pdist.xi_banded(eng, band, params, MODE_AUTO)
    res_b, npall_b, npused_b = pdist.dmat_chunk_banded(eng, band, params, MODE_AUTO, cf.reject,
                                                       hps[0], segments=3)
    if rank == 0:
        cf.fill_neighs(hps)
        want = cf.compute_xi_batch(hps)
        got = full.cpu().numpy()
        assert np.array_equal(got[:, 5].view(np.int64), want[:, 5].view(np.int64))
        for k in range(5):
            np.testing.assert_allclose(got[:, k], want[:, k], rtol=1e-12, atol=1e-300)
        cf.fill_neighs(hps)
        np.random.seed(hps[0])
        one = cf.compute_dmat(hps)
        assert (npall, npused) == (one[6], one[7]) and npused > 500
        assert (npall_b, npused_b) == (one[6], one[7])
        for got_res in (res, res_b):
            for k, (a, b) in enumerate(zip(got_res, one[:6])):
                a = a.cpu().numpy()
                scale = np.abs(b).max()
                assert np.abs(a - b).max() <= 1e-11 * scale, (k, np.abs(a - b).max(), scale)
        got_b = full_b.cpu().numpy()
        assert np.array_equal(got_b[:, 5].view(np.int64), want[:, 5].view(np.int64))
        for k in range(5):
            np.testing.assert_allclose(got_b[:, k], want[:, k], rtol=1e-12, atol=1e-300)
        print("multi-gpu ok: world %d, %d binned pairs, NPALL %d NPUSED %d" % (
            world, int(want[:, 5].view(np.int64).sum()), npall, npused))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
