"""CPU, world_size 2 over gloo: the host-side multi-GPU logic -- LPT partition of HEALPix pixels,
gather of per-pixel blocks to rank 0, sum-reduction of the distortion-matrix accumulators."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from picca_b200 import catalog, dist as pdist
from tests import helpers


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    data, num, z_min, _, cosmo = helpers.small_sample(n=120, seed=7, max_pix=40)
    host = catalog.pack(data)
    work = pdist.estimate_work(host, host, 0.02)
    parts = pdist.lpt_partition(work, world)
    mine = parts[rank]
    nb = 9
    # stand-in for the per-pixel blocks a rank's GPU would produce: a function of the pixel index
    local = torch.stack([torch.full((6, nb), float(k + 1), dtype=torch.float64) for k in mine]) \
        if len(mine) else torch.zeros((0, 6, nb), dtype=torch.float64)
    full = pdist.gather_rows(local, mine, len(host.healpixs))
    dm = torch.full((4, 5), float(rank + 1), dtype=torch.float64)
    pdist.allreduce_dmat([dm])
    if rank == 0:
        np.save(os.path.join(out_dir, "full.npy"), full.numpy())
        np.save(os.path.join(out_dir, "dm.npy"), dm.numpy())
        np.save(os.path.join(out_dir, "loads.npy"),
                np.array([work[p].sum() for p in parts]))
    else:
        assert full is None
    dist.destroy_process_group()


def test_partition_gather_and_reduce_world2(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    full = np.load(tmp_path / "full.npy")
    n_hp = full.shape[0]
    assert np.array_equal(full[:, 0, 0], np.arange(1, n_hp + 1, dtype=np.float64))
    assert np.all(np.load(tmp_path / "dm.npy") == 3.0)
    loads = np.load(tmp_path / "loads.npy")
    assert loads.max() / loads.mean() < 1.35  # LPT keeps the shards balanced


def test_lpt_is_a_partition_and_balanced():
    rng = np.random.default_rng(0)
    work = rng.pareto(1.5, 500) + 0.1
    parts = pdist.lpt_partition(work, 8)
    allidx = np.sort(np.concatenate(parts))
    assert np.array_equal(allidx, np.arange(500))
    loads = np.array([work[p].sum() for p in parts])
    assert loads.max() <= loads.mean() + work.max()


def test_keep_mask_draw_and_shards():
    """The chunk-wide --rej draw equals the reference's per-forest draws after
    np.random.seed(healpixs[0]) (picca_dmat.py:36, cf.py:444); its per-rank shards are disjoint
    and their union is the chunk's mask, so all-reduced sums equal the one-process sums."""
    counts = [3, 0, 7, 1, 12, 5]
    seed, reject = 1234, 0.6
    np.random.seed(seed)
    want = np.concatenate([np.random.rand(n) > reject for n in counts])
    got = pdist.draw_keep_mask(sum(counts), reject, seed)
    assert np.array_equal(got, want)
    pair_row = np.repeat([0, 0, 1, 2, 2, 3], counts)
    owner = np.array([1, 0, 1, 0])
    shards = [pdist.shard_keep_mask(got, pair_row, owner, r) for r in range(2)]
    assert not np.any(shards[0] & shards[1])
    assert np.array_equal(shards[0] | shards[1], got)
    assert np.array_equal(shards[1], got & np.isin(pair_row, [0, 2]))


def _gather_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = torch.arange(3 * rank, 3 * rank + 2 + 3 * rank, dtype=torch.int32)  # lengths 2 and 5
    got = pdist.all_gather_concat(mine)
    if rank == 0:
        np.save(os.path.join(out_dir, "cat.npy"), got.numpy())
    dist.destroy_process_group()


def test_all_gather_concat_of_ragged_tensors(tmp_path):
    """the per-band neighbour counts of the distortion matrix (dist.dmat_chunk_banded) have
    different lengths on every rank"""
    mp.spawn(_gather_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    assert np.array_equal(np.load(tmp_path / "cat.npy"), [0, 1, 3, 4, 5, 6, 7])


def test_band_bounds_and_halo_cover_every_neighbour():
    """Band shards: contiguous, balanced, and the packed rows [h0, h1) of a band contain every
    true neighbour (ang < ang_max, brute force) of the band's forests."""
    from oracle import _host
    from picca_b200 import synth
    data, num, z_min, _, cosmo = synth.make_forests(2500, seed=8, nside=32, ra_deg=(0., 20.),
                                                    dec_deg=(0., 14.), max_pix=12)
    ang_max = synth.compute_ang_max(cosmo, 200., z_min)
    idx = pdist.RowIndex(data)
    work = idx.work(idx, ang_max)
    host = catalog.pack(data)
    np.testing.assert_allclose(work, pdist.estimate_work(host, host, ang_max), rtol=1e-12)
    cat = _host.catalogue(data)
    row_of = np.repeat(np.arange(len(idx.healpixs)), idx.counts)
    for world in (1, 2, 5):
        bounds = pdist.band_bounds(work, world)
        assert bounds[0][0] == 0 and bounds[-1][1] == len(work)
        assert all(a[1] == b[0] and a[1] > a[0] for a, b in zip(bounds[:-1], bounds[1:]))
        loads = np.array([work[a:b].sum() for a, b in bounds])
        assert loads.max() <= loads.mean() + work.max()
        for b0, b1 in bounds:
            reach = np.nonzero(idx.near(idx, np.arange(b0, b1), ang_max).any(axis=0))[0]
            h0, h1 = min(reach.min(), b0), max(reach.max() + 1, b1)
            for f in range(idx.first[b0], idx.first[b1], 23):
                d = cat.objs[f]
                ang = _host.angle_between_many(d, cat)
                nb_ = np.nonzero((ang < ang_max) & (cat.thingid != d.thingid))[0]
                assert np.all((row_of[nb_] >= h0) & (row_of[nb_] < h1))
    # more ranks than rows: trailing bands are empty, nothing is lost
    tiny = pdist.band_bounds(np.ones(3), 5)
    assert tiny[0][0] == 0 and tiny[-1][1] == 3 and sum(b - a for a, b in tiny) == 3


def test_exact_row_work_equals_brute_force_and_balances_bands():
    """RowIndex.work_exact: sum of npix1 * npix2 over the forest pairs a row owns (ang < ang_max,
    ra1 > ra2), block by block == brute force; bands cut by it carry even exact loads."""
    from picca_b200 import synth
    ix = synth.make_forest_index(6000, seed=3, nside=32, ra_deg=(0., 40.), dec_deg=(0., 20.), max_pix=20)
    ang_max = synth.compute_ang_max(ix.cosmo, 200., ix.z_min)
    idx = pdist.RowIndex.from_arrays(ix.healpixs, ix.counts, ix.xyz, ix.npix)
    w = idx.work_exact(ang_max, max_elems=1 << 18)     # several blocks
    ok = (ix.xyz @ ix.xyz.T > np.cos(ang_max)) & (ix.ra[:, None] > ix.ra[None, :])
    per_f = (ok * ix.npix[None, :]).sum(axis=1) * ix.npix
    row_of = np.repeat(np.arange(len(ix.healpixs)), ix.counts)
    np.testing.assert_allclose(w, np.bincount(row_of, weights=per_f), rtol=1e-12)
    import torch
    np.testing.assert_allclose(idx.work_exact(ang_max, torch, torch.device("cpu")), w, rtol=1e-12)
    loads = np.array([w[a:b].sum() for a, b in pdist.band_bounds(w, 4)])
    assert loads.max() <= loads.mean() + w.max()
