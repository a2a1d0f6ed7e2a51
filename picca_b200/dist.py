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


class Shard:
    """This rank's share of the HEALPix rows of catalogue 1 (LPT over estimated pair work,
    identical on every rank) and the index arrays its kernels need."""

    def __init__(self, eng, host1, host2, ang_max, world, rank):
        torch = eng.torch
        self.world, self.rank = world, rank
        self.n_rows_total = len(host1.healpixs)
        self.work = estimate_work(host1, host2, ang_max)
        self.parts = lpt_partition(self.work, world)
        self.mine = self.parts[rank]
        hp_first = host1.arrays["hp_first"]
        parts = [np.arange(hp_first[k], hp_first[k + 1], dtype=np.int32) for k in self.mine]
        self.f1_index = np.concatenate(parts) if parts else np.zeros(0, np.int32)
        self.rows = np.concatenate([np.full(len(p), k, np.int32) for k, p in enumerate(parts)]) \
            if parts else np.zeros(0, np.int32)
        self.row_owner = np.zeros(self.n_rows_total, dtype=np.int64)
        for r, part in enumerate(self.parts):
            self.row_owner[part] = r
        self.d_f1 = torch.as_tensor(self.f1_index, device=eng.device)
        self.d_rows = torch.as_tensor(self.rows, device=eng.device)
        # every line of sight of catalogue 1, catalogue order (the --rej stream order)
        self.d_all_f1 = torch.arange(host1.n_los, dtype=torch.int32, device=eng.device)
        owner_of_f1 = np.repeat(self.row_owner, np.diff(hp_first))
        self.d_f1_is_mine = torch.as_tensor(owner_of_f1 == rank, device=eng.device)


def xi_sharded(eng, dev1, dev2, params, shard, mode, cross_obj=False, gather=True):
    """compute_xi of every HEALPix row, each rank doing its LPT share (neighbour search + pair
    kernel + per-row normalisation); rows gathered to rank 0 ([n_rows_total, 6, nb]; None on the
    other ranks).  No collective on the data path: rows are independent (picca_cf.py:454-473)."""
    pairs = eng.neighbours(dev1, dev2, params, mode, shard.d_f1)
    out = eng.xi(dev1, dev2, params, pairs, shard.d_rows, len(shard.mine), cross_obj=cross_obj,
                 normalise=True)
    if shard.world > 1 and gather:
        return gather_rows(out, shard.mine, shard.n_rows_total)
    return out


def dmat_chunk_sharded(eng, dev1, dev2, params, shard, mode, reject, seed, cross_obj=False,
                       segments=8, group=None):
    """Distortion matrix of ONE reference chunk (all HEALPix rows of catalogue 1, seeded with
    ``seed`` = its first pixel as picca_dmat.py:36 does), the kept forest pairs sharded over the
    ranks by owning HEALPix row.

    The --rej draw is a property of the chunk: one legacy MT19937 stream, ``len(neighbours)``
    numbers per forest in catalogue order (cf.py:444, xcf.py:379), so every rank needs the
    neighbour COUNT of every forest (count-only pass, replicated: milliseconds) but the neighbour
    LISTS of its own forests only.  See ``_dmat_stream`` for the rest.

    Returns ([weights_dmat, dmat, r_par_eff, r_trans_eff, z_eff, weight_eff] device tensors,
    NPALL, NPUSED)."""
    torch = eng.torch
    count_all = eng.neighbour_counts(dev1, dev2, params, mode, shard.d_all_f1)
    off_all = torch.zeros(count_all.numel() + 1, dtype=torch.int64, device=eng.device)
    torch.cumsum(count_all, dim=0, out=off_all[1:])
    pairs = eng.neighbours(dev1, dev2, params, mode, shard.d_f1)
    return _dmat_stream(eng, dev1, dev2, params, pairs, shard.f1_index.astype(np.int64), off_all,
                        reject, seed, cross_obj, segments, shard.world, group)


def _dmat_stream(eng, dev1, dev2, params, pairs, my_f1, off_all, reject, seed, cross_obj,
                 segments, world, group):
    """``pairs``: the neighbour lists of this rank's forests, whose positions in the CHUNK's
    catalogue order are ``my_f1`` (ascending); ``off_all``: device int64 [n_los_chunk + 1], the
    offsets of every forest of the chunk in the --rej stream.  The stream is drawn on the host in
    ``segments`` pieces while the device works on the previous piece: piece k of the mask is
    uploaded, this rank's forest pairs inside it are selected on the device and the kernels run
    on them, accumulating into the same matrices; ONE all-reduce(SUM) of a flat buffer (matrix +
    five vectors) follows."""
    torch = eng.torch
    dev = eng.device
    h_off_all = off_all.cpu().numpy()                     # 0.8 MB per 100k forests
    npall = int(h_off_all[-1])
    n_los = h_off_all.size - 1
    # stream position of each of my pairs: start of its forest's draw + rank inside the forest
    d_my_f1 = torch.from_numpy(np.ascontiguousarray(my_f1)).to(dev)
    f1_cat = d_my_f1[pairs.nb_f1.to(torch.int64)]
    pos = off_all[f1_cat] + (torch.arange(pairs.n_pairs, dtype=torch.int64, device=dev) -
                             pairs.nb_offset[pairs.nb_f1.to(torch.int64)])
    # segments of the stream, cut at forest boundaries
    cuts = np.unique(np.linspace(0, n_los, max(1, segments) + 1).astype(np.int64))
    state = np.random.RandomState(seed)
    outs = eng.dmat_outputs(params)
    keep_dev = torch.zeros(max(pairs.n_pairs, 1), dtype=torch.uint8, device=dev)
    npused = npall_cross = 0
    h_my_off = pairs.host_offset()
    for a, b in zip(cuts[:-1], cuts[1:]):
        s0, s1 = int(h_off_all[a]), int(h_off_all[b])
        if s1 == s0:
            continue
        mask = state.rand(s1 - s0) > reject             # host MT19937, reference stream
        npused += int(mask.sum())
        if cross_obj:
            # xcf.py:379-383: a forest whose draw keeps nothing is skipped BEFORE it is counted
            starts = h_off_all[a:b] - s0
            has = h_off_all[a + 1:b + 1] > h_off_all[a:b]
            kept_any = np.zeros(b - a, dtype=bool)
            if has.any():
                kept_any[has] = np.add.reduceat(mask.astype(np.int64), starts[has]) > 0
            npall_cross += int((h_off_all[a + 1:b + 1] - h_off_all[a:b])[kept_any].sum())
        lo, hi = np.searchsorted(my_f1, [a, b])           # my forests inside this segment
        if hi == lo:
            continue
        p0, p1 = int(h_my_off[lo]), int(h_my_off[hi])
        if p1 == p0:
            continue
        d_mask = torch.from_numpy(mask.view(np.uint8)).to(dev, non_blocking=True)
        keep_dev.zero_()
        keep_dev[p0:p1] = d_mask[pos[p0:p1] - s0]
        pairs.nb_keep = keep_dev
        eng.dmat(dev1, dev2, params, pairs, cross_obj=cross_obj, out=outs)
    # one collective over a flat buffer
    if world > 1:
        import torch.distributed as dist
        flat = torch.cat([t.reshape(-1) for t in outs])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        k = 0
        for t in outs:
            t.copy_(flat[k:k + t.numel()].view_as(t))
            k += t.numel()
    return list(outs), (npall_cross if cross_obj else npall), npused


# ---------------------------------------------------------------------------------------------
# Band shards: every rank holds only a contiguous band of HEALPix rows plus the halo its
# neighbour searches reach into -- the host packs and uploads 1/N of the catalogue (+ halo)
# instead of all of it, which is what the end-to-end path costs per rank.
# ---------------------------------------------------------------------------------------------
class RowIndex:
    """Per HEALPix row of a ``data`` dict (ascending pixel id): forests, spectral pixels and the
    bounding cap of the members -- what the partition and the halo need, without packing."""

    def __init__(self, data):
        from . import forest as _forest
        healpixs = sorted(data)
        counts = np.array([len(data[hp]) for hp in healpixs], dtype=np.int64)
        reg = _forest.soa_of(data)
        if reg is not None and reg["objs"] and _forest.registered_clean(data, reg):
            los = reg["los"]
            xyz = np.stack([np.array(los[k], dtype=np.float64) for k in ("x_cart", "y_cart", "z_cart")],
                           axis=1)
            npix = np.diff(np.asarray(reg["offset"], dtype=np.int64)).astype(np.float64)
        else:
            objs = [o for hp in healpixs for o in data[hp]]
            xyz = np.array([[o.x_cart, o.y_cart, o.z_cart] for o in objs], dtype=np.float64)
            npix = np.array([np.size(o.weights) for o in objs], dtype=np.float64)
        self._set(healpixs, counts, xyz, npix)

    @classmethod
    def from_arrays(cls, healpixs, counts, xyz, npix):
        """From per-forest arrays in catalogue order (a survey index that holds no pixel data,
        ``synth.make_forest_index``): ``healpixs`` ascending, ``counts`` forests per row."""
        self = cls.__new__(cls)
        self._set(list(healpixs), np.asarray(counts, dtype=np.int64),
                  np.asarray(xyz, dtype=np.float64).reshape(-1, 3), np.asarray(npix, dtype=np.float64))
        return self

    def _set(self, healpixs, counts, xyz, npix):
        self.healpixs = healpixs
        self.counts = counts
        self.xyz_los, self.npix_los = xyz, npix   # per line of sight, catalogue order
        self.first = np.zeros(len(self.healpixs) + 1, dtype=np.int64)
        np.cumsum(self.counts, out=self.first[1:])
        n = len(self.healpixs)
        if not npix.size:
            self.npix, self.cap, self.cap_rad = np.zeros(n), np.zeros((n, 3)), np.zeros(n)
            return
        # (rows are never empty: a HEALPix pixel is a key of `data` because a forest fell in it)
        start = self.first[:-1]
        self.npix = np.add.reduceat(npix, start)
        c = np.add.reduceat(xyz, start, axis=0)
        norm = np.sqrt((c * c).sum(axis=1))
        c = np.where(norm[:, None] > 0, c / np.where(norm > 0, norm, 1.)[:, None], xyz[start])
        row = np.repeat(np.arange(n), self.counts)
        dots = np.clip((xyz * c[row]).sum(axis=1), -1., 1.)
        self.cap = c
        self.cap_rad = np.arccos(np.minimum.reduceat(dots, start)) + 1e-7

    def near(self, other, rows, ang_max):
        """bool [len(rows), n_other]: can a member of row r be within ang_max of a member of a row
        of ``other`` (bounding caps: a superset, like the device neighbour search)"""
        ang = np.arccos(np.clip(self.cap[rows] @ other.cap.T, -1., 1.))
        return ang <= (ang_max + self.cap_rad[rows, None] + other.cap_rad[None, :])

    def work_exact(self, ang_max, torch=None, device=None, max_elems=1 << 28):
        """Work of every row in the auto-correlation, from the forest pairs themselves: the sum of
        npix1 * npix2 over the pairs (f1 in the row, ang < ang_max, ra1 > ra2: the pairs
        cf.fill_neighs gives the row, cf.py:109-135).  The bounding-cap estimate of ``work`` is off
        by up to 9 % per band on a 1M-forest survey (a ring of HEALPix pixels counts as near as soon
        as two caps touch); bands cut by this weight are even to 0.2 %.  Dot products of the
        unit vectors block by block: on ``device`` with torch (~0.1 s for 1M forests on a B200),
        with NumPy otherwise."""
        n = len(self.healpixs)
        out = np.zeros(n)
        if not self.npix_los.size:
            return out
        xyz, npix = self.xyz_los, self.npix_los
        ra = np.arctan2(xyz[:, 1], xyz[:, 0]) % (2. * np.pi)     # ordering only
        cos_max = float(np.cos(ang_max))
        on_dev = torch is not None and device is not None
        if on_dev:
            X = torch.as_tensor(xyz, dtype=torch.float64, device=device)
            R = torch.as_tensor(ra, dtype=torch.float64, device=device)
            W = torch.as_tensor(npix, dtype=torch.float64, device=device)
            out_t = torch.zeros(n, dtype=torch.float64, device=device)
        r0 = 0
        while r0 < n:
            # rows [r0, r1) against the forests of every row near one of them
            r1 = r0 + 1
            near = self.near(self, np.arange(r0, min(n, r0 + 256)), ang_max)
            reach = near[0]
            while r1 < n and r1 - r0 < 256:
                grown = reach | near[r1 - r0]
                n_c = int(self.counts[grown].sum())
                if (self.first[r1 + 1] - self.first[r0]) * n_c > max_elems:
                    break
                reach = grown
                r1 += 1
            rows_c = np.nonzero(reach)[0]
            cand = np.concatenate([np.arange(self.first[q], self.first[q + 1]) for q in rows_c])
            a, b = int(self.first[r0]), int(self.first[r1])
            row_of = np.repeat(np.arange(r1 - r0), self.counts[r0:r1])
            if on_dev:
                ci = torch.as_tensor(cand, device=device)
                ok = (X[a:b] @ X[ci].T > cos_max) & (R[a:b, None] > R[ci][None, :])
                per_f = (ok * W[ci][None, :]).sum(dim=1) * W[a:b]
                out_t[r0:r1].index_add_(0, torch.as_tensor(row_of, device=device), per_f)
            else:
                ok = (xyz[a:b] @ xyz[cand].T > cos_max) & (ra[a:b, None] > ra[None, cand])
                per_f = (ok * npix[None, cand]).sum(axis=1) * npix[a:b]
                out[r0:r1] = np.bincount(row_of, weights=per_f, minlength=r1 - r0)
            r0 = r1
        return out_t.cpu().numpy() if on_dev else out

    def work(self, other, ang_max):
        out = np.zeros(len(self.healpixs))
        for a in range(0, len(out), 512):
            rows = np.arange(a, min(a + 512, len(out)))
            out[rows] = self.npix[rows] * (self.near(other, rows, ang_max) * other.npix[None, :]).sum(axis=1)
        return out


def band_bounds(work, n_parts):
    """Contiguous bands [b0, b1) of the rows with balanced cumulative work (every band non-empty
    when there are at least ``n_parts`` rows)."""
    work = np.asarray(work, dtype=np.float64)
    n = work.size
    c = np.concatenate([[0.], np.cumsum(work)])
    cuts = [0]
    for k in range(1, n_parts):
        b = int(np.searchsorted(c, c[-1] * k / n_parts, side="left"))
        b = min(max(b, cuts[-1] + 1), n - (n_parts - k)) if n >= n_parts else min(b, n)
        cuts.append(max(b, cuts[-1]))
    cuts.append(n)
    return [(cuts[k], cuts[k + 1]) for k in range(n_parts)]


def all_gather_concat(t, group=None):
    """Concatenation over the ranks, in rank order, of 1-D tensors of different lengths."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    lens = [torch.zeros(1, dtype=torch.int64, device=t.device) for _ in range(world)]
    dist.all_gather(lens, torch.tensor([t.numel()], dtype=torch.int64, device=t.device), group=group)
    lens = [int(x.item()) for x in lens]
    cap = max(max(lens), 1)
    pad = torch.zeros(cap, dtype=t.dtype, device=t.device)
    pad[:t.numel()] = t
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([b[:n] for b, n in zip(bufs, lens)])


class BandShard:
    """Rank ``rank`` of ``world``: the band [b0, b1) of HEALPix rows it owns and, packed and
    resident in HBM, the rows [h0, h1) >= band that contain every possible neighbour of the
    band's forests (auto-correlation: ``data`` against itself)."""

    def __init__(self, eng, data, ang_max, world, rank, ang_correlation=False, index=None,
                 band_source=None):
        """``data``: the whole catalogue (registered SoA: the band is sliced out of it) -- or
        None with ``index`` (a ``RowIndex`` of the whole survey) and ``band_source(h0, h1)``
        returning the ``data`` dict of the rows [h0, h1) only: the rank then never holds (or
        generates, or reads) more than its band + halo."""
        from . import catalog as _catalog
        torch = eng.torch
        self.world, self.rank = world, rank
        idx = RowIndex(data) if index is None else index
        self.n_rows_total = len(idx.healpixs)
        self.healpixs = idx.healpixs
        # bands of equal work: exact forest-pair weights when there is more than one band (one
        # pass of dot products on the device; rank 0's result is everybody's, so that no two
        # ranks can ever cut the rows differently)
        if world > 1:
            work = torch.as_tensor(idx.work_exact(ang_max, torch, eng.device) if rank == 0 else
                                   np.zeros(len(idx.healpixs)), device=eng.device)
            import torch.distributed as tdist
            if tdist.is_available() and tdist.is_initialized():
                tdist.broadcast(work, src=0)
                work = work.cpu().numpy()
            else:   # ranks emulated one after the other in one process (tests)
                work = idx.work_exact(ang_max, torch, eng.device)
        else:
            work = idx.work(idx, ang_max)
        self.work = work
        self.bounds = band_bounds(work, world)
        self.b0, self.b1 = self.bounds[rank]
        if self.b1 > self.b0:
            reach = np.nonzero(idx.near(idx, np.arange(self.b0, self.b1), ang_max).any(axis=0))[0]
            self.h0, self.h1 = int(min(reach.min(), self.b0)), int(max(reach.max() + 1, self.b1))
        else:
            self.h0, self.h1 = self.b0, self.b1
        if index is None:
            self.data = data
            self.host = _catalog.pack(data, ang_correlation=ang_correlation, defer_products=True,
                                      rows=(self.h0, self.h1))
        else:
            self.data = band_source(self.h0, self.h1)
            assert sorted(self.data) == list(idx.healpixs[self.h0:self.h1])
            self.host = _catalog.pack(self.data, ang_correlation=ang_correlation,
                                      defer_products=True)
        self.dev = eng.device_catalog(self.host, cache=False)
        self.mine = np.arange(self.b0, self.b1, dtype=np.int64)       # global rows of the band
        first = self.host.arrays["hp_first"]
        lo, hi = self.b0 - self.h0, self.b1 - self.h0                  # the band inside the halo
        self.f1_index = np.arange(first[lo], first[hi], dtype=np.int32)
        self.rows = (np.repeat(np.arange(lo, hi, dtype=np.int32), np.diff(first[lo:hi + 1])) - lo
                     ).astype(np.int32)
        self.d_f1 = torch.as_tensor(self.f1_index, device=eng.device)
        self.d_rows = torch.as_tensor(self.rows, device=eng.device)
        # positions of the band's forests in the whole catalogue (the --rej stream order)
        self.global_f1 = int(idx.first[self.b0]) + np.arange(self.f1_index.size, dtype=np.int64)
        self.n_los_total = int(idx.first[-1])
        self.h2d_bytes = int(self.host.nbytes())


def xi_banded(eng, shard, params, mode, gather=True):
    """compute_xi of every HEALPix row from band shards: neighbour search + pair kernel + per-row
    normalisation on the band, rows gathered to rank 0 ([n_rows_total, 6, nb]; None elsewhere)."""
    pairs = eng.neighbours(shard.dev, shard.dev, params, mode, shard.d_f1)
    out = eng.xi(shard.dev, shard.dev, params, pairs, shard.d_rows, len(shard.mine), normalise=True)
    if shard.world > 1 and gather:
        return gather_rows(out, shard.mine, shard.n_rows_total)
    return out


def dmat_chunk_banded(eng, shard, params, mode, reject, seed, segments=8, group=None):
    """``dmat_chunk_sharded`` from band shards: a rank can only count the neighbours of ITS
    forests, so the counts of the whole chunk -- the offsets of the --rej stream -- are
    all-gathered (bands are contiguous and ordered by rank: a concatenation)."""
    torch = eng.torch
    count_mine = eng.neighbour_counts(shard.dev, shard.dev, params, mode, shard.d_f1)
    count_all = all_gather_concat(count_mine, group) if shard.world > 1 else count_mine
    assert count_all.numel() == shard.n_los_total
    off_all = torch.zeros(count_all.numel() + 1, dtype=torch.int64, device=eng.device)
    torch.cumsum(count_all, dim=0, out=off_all[1:])
    pairs = eng.neighbours(shard.dev, shard.dev, params, mode, shard.d_f1)
    return _dmat_stream(eng, shard.dev, shard.dev, params, pairs, shard.global_f1, off_all, reject,
                        seed, False, segments, shard.world, group)
