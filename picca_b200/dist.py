"""Multi-GPU sharding of the pair-counting path: one process per GPU, HEALPix pixels partitioned
by estimated pair work (longest-processing-time first), per-pixel blocks gathered to rank 0 and
the distortion matrix summed with one NCCL all-reduce.

The reference's only parallelism is a fork pool over HEALPix pixels whose results are stacked /
summed on the host (picca_cf.py:454-473, picca_dmat.py:471-501); rows of compute_xi are
independent, so no collective is needed on the data path of the correlation function.
"""
import numpy as np


def estimate_work(host_cat, host_cat2, ang_max):
    """Estimated candidate pixel pairs per HEALPix pixel of catalogue 1: pixels of the row's
    forests times pixels of all forests in pixels whose bounding caps can be within ang_max."""
    A, B = host_cat.arrays, host_cat2.arrays
    npix1 = np.diff(A["offset"]).astype(np.float64)
    npix2 = np.diff(B["offset"]).astype(np.float64)
    tot1 = np.add.reduceat(npix1, A["hp_first"][:-1]) if host_cat.n_los else np.zeros(0)
    tot2 = np.add.reduceat(npix2, B["hp_first"][:-1]) if host_cat2.n_los else np.zeros(0)
    c1 = np.stack([A["cap_x"], A["cap_y"], A["cap_z"]], axis=1)
    c2 = np.stack([B["cap_x"], B["cap_y"], B["cap_z"]], axis=1)
    work = np.zeros(len(tot1))
    step = 512
    for a in range(0, len(tot1), step):
        ang = np.arccos(np.clip(c1[a:a + step] @ c2.T, -1., 1.))
        near = ang <= (ang_max + A["cap_rad"][a:a + step, None] + B["cap_rad"][None, :])
        work[a:a + step] = tot1[a:a + step] * (near * tot2[None, :]).sum(axis=1)
    return work


def lpt_partition(work, n_parts):
    """Longest-processing-time-first assignment.  Returns a list of index arrays (ascending)."""
    order = np.argsort(-np.asarray(work, dtype=np.float64), kind="stable")
    loads = np.zeros(n_parts)
    parts = [[] for _ in range(n_parts)]
    for k in order:
        p = int(np.argmin(loads))
        parts[p].append(int(k))
        loads[p] += work[k]
    return [np.array(sorted(p), dtype=np.int64) for p in parts]


def gather_rows(local_rows, local_index, n_rows_total, group=None, dst=0):
    """Gather per-HEALPix blocks [n_local, 6, nb] computed by each rank into the full
    [n_rows_total, 6, nb] array on ``dst`` (other ranks get None).  Works with NCCL (device
    tensors) and gloo (CPU tensors)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    counts = [torch.zeros(1, dtype=torch.int64, device=local_rows.device) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([local_rows.shape[0]], dtype=torch.int64,
                                         device=local_rows.device), group=group)
    counts = [int(c.item()) for c in counts]
    cap = max(counts) if counts else 0
    shape = (cap,) + tuple(local_rows.shape[1:])
    pad = torch.zeros(shape, dtype=local_rows.dtype, device=local_rows.device)
    pad[:local_rows.shape[0]] = local_rows
    idx = torch.full((cap,), -1, dtype=torch.int64, device=local_rows.device)
    idx[:local_rows.shape[0]] = torch.as_tensor(local_index, dtype=torch.int64,
                                                device=local_rows.device)
    bufs = [torch.empty_like(pad) for _ in range(world)]
    ibufs = [torch.empty_like(idx) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    dist.all_gather(ibufs, idx, group=group)
    if rank != dst:
        return None
    full = torch.zeros((n_rows_total,) + tuple(local_rows.shape[1:]), dtype=local_rows.dtype,
                       device=local_rows.device)
    for r in range(world):
        n = counts[r]
        if n:
            full[ibufs[r][:n]] = bufs[r][:n]
    return full


def allreduce_dmat(tensors, group=None):
    """Sum the distortion-matrix accumulators of all ranks in place (the NumPy ``.sum(axis=0)``
    over workers of picca_dmat.py:494-501, as one collective per tensor)."""
    import torch.distributed as dist
    for t in tensors:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return tensors


def draw_keep_mask(n_pairs, reject, seed):
    """The ``--rej`` draw of one reference chunk: ``np.random.seed(healpixs[0])``
    (picca_dmat.py:36) followed by ``rand(len(neighbours)) > reject`` per forest in catalogue
    order (cf.py:444).  Consecutive ``rand(n)`` calls consume the legacy MT19937 stream exactly
    like one ``rand(sum n)``, so one call reproduces the whole chunk; every rank draws the same
    mask from the same seed."""
    state = np.random.RandomState(seed)
    return state.rand(int(n_pairs)) > reject


def shard_keep_mask(keep, pair_row, row_owner, rank):
    """Restrict the chunk-wide keep mask to the forest pairs whose owning HEALPix row belongs to
    ``rank`` (``row_owner[row]`` = rank).  The union over ranks is the chunk's mask and the
    shards are disjoint, so the all-reduced distortion matrix equals the single-process one."""
    return keep & (np.asarray(row_owner)[np.asarray(pair_row)] == rank)


def dmat_sharded(eng, dev1, dev2, params, pairs, keep_local, group=None, world=1):
    """Distortion matrix of this rank's share of the kept forest pairs, then ONE all-reduce(SUM)
    per accumulator across ranks (the NumPy ``.sum(axis=0)`` of picca_dmat.py:494-501)."""
    torch = eng.torch
    pairs.nb_keep = torch.from_numpy(np.ascontiguousarray(keep_local, dtype=np.uint8)).to(eng.device)
    res = list(eng.dmat(dev1, dev2, params, pairs))
    if world > 1:
        allreduce_dmat(res, group=group)
    return res
