#!/usr/bin/env python
"""Benchmark of the forest pair-counting hot path (BASELINE.json metric: binned forest-pixel
pairs per second, auto-correlation, synthetic 100k-forest DR16-like sample, nside 32).

    python bench.py --gpus N --steps K --warmup W            # CUDA path (picca_b200)
    python bench.py --impl reference --steps K --warmup W    # CPU arm: the reference's Numba path

CUDA arm.  A step is one pass of the hot path over the workload: device neighbour search
(fill_neighs) + pair kernel (compute_xi) + normalisation, per-HEALPix blocks for every pixel.
`value`: catalogue resident in HBM.  `e2e`: the same pass entered through the plugin API with HOST
data -- ``cf.fill_neighs(healpixs)`` + ``cf.compute_xi_batch(healpixs)`` on the ``data`` dict, i.e.
catalogue packing on the host, H2D, record packing, neighbour search, kernel, D2H of the blocks,
all inside the timed region, every step.  N > 1: HEALPix rows LPT-sharded over the ranks, blocks
gathered to rank 0 inside the timed region, time = max over ranks.  Further legs on the same
sample: the distortion matrix (BASELINE config 4, --rej 0.99) and the forest x quasar
cross-correlation (config 3 at its surface density on this footprint).  `parity_check`: rows of
the timed GPU result against the oracle (C restatement of the Numba kernel), at every N.

Reference arm / cpu_baseline.  The UNMODIFIED reference (staged under baseline/_ref by
scripts/stage_reference.py; healpy / fitsio stand-ins from tests/refharness) runs its own
``cf.fill_neighs`` + ``cf.compute_xi`` (Numba) in a fork pool over all host cores, exactly the call
pattern of picca_cf.py:449-463, on a bounded sample of the same workload: each step takes
4 x cores tasks of a few forests each from HEALPix pixels spread over the footprint (other
pixels every step).  If the staged reference or numba is missing the oracle port is timed
instead and the line says kind "port".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[1] / SURVEY.md 8d C2
    "c2_100k": dict(n_forest=100000, seed=20260102, nside=32, ra_deg=(0., 120.),
                    dec_deg=(0., 40.2)),
    # reduced footprints at the same surface density and forest shape (development only)
    "c2_20k": dict(n_forest=20000, seed=20260102, nside=32, ra_deg=(0., 53.6),
                   dec_deg=(0., 18.)),
    "c2_5k": dict(n_forest=5000, seed=20260102, nside=32, ra_deg=(0., 26.8), dec_deg=(0., 9.)),
    # BASELINE.json configs[2] / SURVEY.md 8d C3: 300k forests over ~14 000 deg^2 (with 500k quasars
    # on the same footprint, scripts/perf_c3_xcf.py)
    "c3_300k": dict(n_forest=300000, seed=20260103, nside=32, ra_deg=(0., 360.),
                    dec_deg=(0., 42.8)),
    # one eighth of BASELINE.json configs[4] / SURVEY.md 8d C5 (1M forests over ~14 000 deg^2,
    # ~71 per deg^2): the share one GPU of an 8-GPU box holds, at the full surface density
    "c5_eighth": dict(n_forest=125000, seed=20260105, nside=32, ra_deg=(0., 60.),
                      dec_deg=(0., 30.8)),
    # BASELINE.json configs[4] / SURVEY.md 8d C5 at full size
    "c5_1m": dict(n_forest=1000000, seed=20260105, nside=32, ra_deg=(0., 360.),
                  dec_deg=(0., 42.8)),
}
CF_CFG = dict(num_bins_r_par=50, num_bins_r_trans=50, r_par_max=200., r_par_min=0.,
              r_trans_max=200., nside=32)
# picca_xcf.py defaults (:107-112): np = 100 over r_par in [-200, 200], nt = 50
XCF_CFG = dict(num_bins_r_par=100, num_bins_r_trans=50, r_par_max=200., r_par_min=-200.,
               r_trans_max=200., nside=32, alpha_obj=1.44)
QUASARS_PER_FOREST = 500000. / 300000.   # BASELINE.json configs[2]: 500k quasars per 300k forests
DMAT_REJECT = 0.99  # BASELINE.json configs[3]
FLOPS_PER_PAIR = 30.      # algorithmic FP64 ops per binned pair (cf)
FLOPS_PER_PAIR_XCF = 28.


def make_workload(name):
    from picca_b200 import synth   # NumPy generator only: no CUDA library is loaded by it
    kw = dict(WORKLOADS[name])
    n = kw.pop("n_forest")
    data, num, z_min, z_max, cosmo = synth.make_forests(n, **kw)
    ang_max = synth.compute_ang_max(cosmo, CF_CFG["r_trans_max"], z_min)
    return data, num, ang_max, cosmo, z_min


def make_quasars(name, cosmo):
    from picca_b200 import synth
    kw = WORKLOADS[name]
    nq = int(round(kw["n_forest"] * QUASARS_PER_FOREST))
    return synth.make_quasars(nq, seed=20260103, nside=kw["nside"], ra_deg=kw["ra_deg"],
                              dec_deg=kw["dec_deg"], cosmo=cosmo), nq


class Cfg:
    """stand-in for the module globals of picca.cf (what picca_cf.py assigns, :343-368)"""


def configure(mod, data, num, ang_max, **over):
    from tests import helpers
    helpers.configure(mod, data, num, ang_max, **dict(CF_CFG, **over))


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True,
                                     text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(r[2 + k] == "Active" for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(self.rows[0][1]) if self.rows[0][1].replace(".", "").isdigit()
                else None, "reasons": reasons, "samples": len(self.rows)}


# --------------------------------------------------------------------------------------------
# CPU side: samples of the workload, the reference's Numba path, the oracle port
# --------------------------------------------------------------------------------------------
def spread_pixels(healpixs, count, step=0):
    """``count`` HEALPix pixels spread over the footprint: a fixed seeded permutation of the
    occupied pixels, consumed ``count`` at a time (other pixels every step)."""
    perm = np.random.default_rng(20261017).permutation(len(healpixs))
    return [healpixs[perm[(step * count + k) % len(healpixs)]] for k in range(count)]


def sample_tasks(data, healpixs, n_tasks, forests_per_task, step):
    """Tasks of a few forests each: task t = the first ``forests_per_task`` forests of the t-th
    sampled pixel.  Finer than the reference's one-pixel tasks so that a bounded sample still
    loads every core; the pairs of a task are those its forests own against the WHOLE catalogue."""
    return [(hp, list(range(min(forests_per_task, len(data[hp])))))
            for hp in spread_pixels(healpixs, n_tasks, step)]


class OracleSoA:
    """The catalogue as the oracle's batch driver wants it (CSR + one array per field), built
    with NumPy from the data dict -- no product code involved."""

    def __init__(self, data):
        from oracle import _host
        self.cat = _host.catalogue(data)
        objs = self.cat.objs
        npix = np.array([len(o.weights) for o in objs], dtype=np.int64)
        self.offset = np.zeros(len(objs) + 1, dtype=np.int64)
        np.cumsum(npix, out=self.offset[1:])
        cat = lambda name: np.ascontiguousarray(np.concatenate([getattr(o, name) for o in objs]))
        self.z, self.r_comov, self.dist_m = cat("z"), cat("r_comov"), cat("dist_m")
        self.weights, self.delta = cat("weights"), cat("delta")
        self.z_qso = np.array([o.z_qso for o in objs], dtype=np.float64)

    def index_of(self, hp, k):
        return self.cat.first[hp] + k


def port_run_tasks(soa, cfg, ang_max, tasks, threads):
    """Oracle port (C restatement of cf.compute_xi_forest_pairs_fast, pthreads over forests) on
    ``tasks``: neighbour search with NumPy as cf.fill_neighs does (cf.py:109-135), then the pair
    loops.  Returns (rows [n_tasks, 6, nb] un-normalised, seconds incl. neighbour search)."""
    import ctypes
    from oracle import _host, _kernels
    cat = soa.cat
    t0 = time.perf_counter()
    f1_index, rows, nb_off, nb_idx, nb_ang = [], [], [0], [], []
    for r, (hp, ks) in enumerate(tasks):
        for k in ks:
            f1 = soa.index_of(hp, k)
            d = cat.objs[f1]
            ang = _host.angle_between_many(d, cat)
            w = (cat.thingid != d.thingid) & (ang < ang_max) & (d.ra > cat.ra)  # cf.py:109-135
            idx = np.nonzero(w)[0]
            f1_index.append(f1)
            rows.append(r)
            nb_idx.append(idx)
            nb_ang.append(ang[idx])
            nb_off.append(nb_off[-1] + idx.size)
    p = _kernels.params_from_module(cfg)
    nb = p.num_bins_r_par * p.num_bins_r_trans
    out = np.zeros((len(tasks), 6, nb))
    f1_index = np.array(f1_index, dtype=np.int64)
    rows = np.array(rows, dtype=np.int64)
    nb_off = np.array(nb_off, dtype=np.int64)
    nb_idx = np.concatenate(nb_idx).astype(np.int64)
    nb_ang = np.concatenate(nb_ang).astype(np.float64)
    lib = _kernels.lib()
    dp, lp = _kernels.dp, _kernels.lp
    lib.orc_xi_auto_batch(
        ctypes.byref(p), lp(soa.offset), dp(soa.z), dp(soa.r_comov), dp(soa.dist_m),
        dp(soa.weights), dp(soa.delta), dp(soa.z_qso), lp(soa.offset), dp(soa.z),
        dp(soa.r_comov), dp(soa.dist_m), dp(soa.weights), dp(soa.delta), dp(soa.z_qso),
        ctypes.c_int64(len(f1_index)), lp(f1_index), lp(rows), lp(nb_off), lp(nb_idx),
        dp(nb_ang), None, ctypes.c_int64(len(tasks)), dp(out), ctypes.c_int32(threads))
    return out, time.perf_counter() - t0


_REF = {}


def _ref_corr_func(healpixs):
    """verbatim body of picca_cf.py:21-37 (corr_func), bound to the live reference module"""
    cf = _REF["cf"]
    cf.fill_neighs(healpixs)
    return cf.compute_xi(healpixs)


class NumbaReference:
    """The unmodified reference's picca.cf (Numba) from baseline/_ref, fed the workload as its
    own ``picca.data.Delta`` objects (array views, no copies)."""

    def __init__(self, data, num, ang_max):
        os.environ["PICCA_B200_FORCE_STAGED_REF"] = "1"   # never /root/reference from bench.py
        from tests.refharness import shims
        if not shims.reference_available():
            raise ImportError("reference tree not staged (baseline/_ref)")
        import numba  # noqa: F401  (the reference's kernels are @njit)
        assert shims.install()
        import warnings
        warnings.filterwarnings("ignore")
        import picca.cf
        from picca.data import Delta
        self.cf = cf = picca.cf
        cf.userprint = lambda *a, **k: None
        self.data = {}
        for hp, forests in data.items():
            out = []
            for d in forests:
                r = Delta(d.thingid, d.ra, d.dec, d.z_qso, d.plate, d.mjd, d.fiberid,
                          d.log_lambda, d.weights, None, d.delta, d.order, None, None, None,
                          None, None)
                r.z, r.r_comov, r.dist_m = d.z, d.r_comov, d.dist_m
                out.append(r)
            self.data[hp] = out
        configure(cf, self.data, num, ang_max)
        _REF["cf"] = cf
        self.path = os.path.dirname(picca.cf.__file__)

    def warm_jit(self):
        """Compile the @njit kernel in the parent (excluded from every timed region):
        forked workers inherit the compiled code."""
        from multiprocessing import Lock, Value
        cf = self.cf
        cf.counter, cf.lock = Value("i", 0), Lock()
        hp = sorted(self.data)[0]
        cf.data[-1] = cf.data[hp][:1]
        _ref_corr_func([-1])
        del cf.data[-1]

    def run_tasks(self, tasks, nproc):
        """picca_cf.py:449-463: counter + lock, a FORK pool of ``nproc`` workers, one task per
        ``corr_func`` call, timed like the script's own "Time computing correlation function".
        A task is entered as an extra key of ``cf.data`` (negative ids never come out of
        query_disc, so the neighbour search still sees exactly the real catalogue)."""
        import multiprocessing
        from multiprocessing import Lock, Value
        cf = self.cf
        keys = []
        for t, (hp, ks) in enumerate(tasks):
            cf.data[-(t + 1)] = [self.data[hp][k] for k in ks]
            keys.append([-(t + 1)])
        cf.counter, cf.lock = Value("i", 0), Lock()
        t1 = time.time()
        context = multiprocessing.get_context("fork")
        pool = context.Pool(processes=nproc)
        res = pool.map(_ref_corr_func, keys)
        pool.close()
        t2 = time.time()
        pool.join()
        for key in keys:
            del cf.data[key[0]]
        rows = np.array([np.stack([np.asarray(r[k], dtype=np.float64) for k in range(5)] +
                                  [np.asarray(r[5], dtype=np.int64).view(np.float64)])
                         for r in res])
        return rows, t2 - t1


def cpu_reference(data, num, ang_max):
    """(runner, kind, note): the Numba reference when it is staged and importable, else the
    oracle port."""
    try:
        ref = NumbaReference(data, num, ang_max)
        ref.warm_jit()
        return ref, "reference", "unmodified picca.cf (Numba) from %s, fork pool" % os.path.relpath(
            ref.path, ROOT)
    except Exception as err:  # no staged reference / numba on this box
        return None, "port", "oracle C port (%s: %s)" % (type(err).__name__, err)


def cpu_timed_steps(data, num, ang_max, steps, warmup, tasks_per_core, forests_per_task):
    """K timed steps (after W untimed ones) of the CPU arm.  Returns the JSON fragment."""
    cores = os.cpu_count() or 1
    healpixs = sorted(data)
    ref, kind, note = cpu_reference(data, num, ang_max)
    if ref is None:
        soa = OracleSoA(data)
        cfg = Cfg()
        configure(cfg, data, num, ang_max)
    n_tasks = tasks_per_core * cores
    times, pairs, forests = [], 0, 0
    for step in range(warmup + steps):
        tasks = sample_tasks(data, healpixs, n_tasks, forests_per_task, step)
        if ref is not None:
            rows, dt = ref.run_tasks(tasks, cores)
        else:
            rows, dt = port_run_tasks(soa, cfg, ang_max, tasks, cores)
        if step >= warmup:
            times.append(dt)
            pairs += int(rows[:, 5].view(np.int64).sum())
            forests += sum(len(ks) for _, ks in tasks)
    total = float(np.sum(times))
    desc = ("%d tasks x <=%d forests per step from %d HEALPix pixels spread over the footprint "
            "(other pixels every step): %d forests, %d binned pairs in %d timed steps; %s"
            % (n_tasks, forests_per_task, n_tasks, forests, pairs, steps, note))
    return {"value": pairs / total, "unit": "pairs/s", "cores": cores, "kind": kind,
            "sample": desc, "ms_per_step": 1e3 * total / max(1, steps)}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation on every host core."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    data, num, ang_max, _, _ = make_workload(args.workload)
    frag = cpu_timed_steps(data, num, ang_max, args.steps, args.warmup, args.cpu_tasks_per_core,
                           args.cpu_forests_per_task)
    line = {
        "impl": "reference", "metric": "binned forest-pixel pairs/sec (cf auto-correlation)",
        "value": frag["value"], "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": frag["ms_per_step"], "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "np": 50, "nt": 50, "rp_max": 200., "rt_max": 200.,
                   "nside": 32},
        "cpu_baseline": {k: frag[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": frag["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def parity_check(data, num, ang_max, healpixs, gpu_rows, n_pixels, threads, chosen=None):
    """Whole HEALPix rows of the timed GPU result against the oracle port on the same pixels:
    num_pairs bit for bit, the five fp64 sums to 1e-9 (north_star).  ``chosen``: the pixels to
    check (default: ``n_pixels`` spread over the footprint)."""
    soa = OracleSoA(data)
    cfg = Cfg()
    configure(cfg, data, num, ang_max)
    if chosen is None:
        chosen = spread_pixels(healpixs, n_pixels, step=7)
    tasks = [(hp, list(range(len(data[hp])))) for hp in chosen]
    want, _ = port_run_tasks(soa, cfg, ang_max, tasks, threads)
    w = want[:, 0] > 0
    for k in (1, 2, 3, 4):   # the per-call normalisation of cf.py:242-246
        want[:, k][w] /= want[:, 0][w]
    got = np.stack([gpu_rows[healpixs.index(hp)] for hp in chosen])
    counts_equal = bool(np.array_equal(got[:, 5].view(np.int64), want[:, 5].view(np.int64)))
    worst = 0.
    for k in range(5):
        scale = np.maximum(np.abs(want[:, k]), 1e-300)
        if k == 1:  # xi: a sum of signed terms, compared on the scale of the histogram
            scale = np.maximum(scale, 1e-3 * np.abs(want[:, k]).max())
        worst = max(worst, float((np.abs(got[:, k] - want[:, k]) / scale).max()))
    return {"checker": "oracle C port (oracle/picca_oracle.c)", "healpix": [int(h) for h in chosen],
            "forests": int(sum(len(t[1]) for t in tasks)),
            "binned_pairs": int(want[:, 5].view(np.int64).sum()),
            "num_pairs_equal": counts_equal, "max_rel_err": worst, "tolerance": 1e-9,
            "ok": bool(counts_equal and worst <= 1e-9)}


# --------------------------------------------------------------------------------------------
# CUDA arm
# --------------------------------------------------------------------------------------------
def load_profile_json(name):
    try:
        with open(os.path.join(ROOT, "profiles", name)) as f:
            return json.load(f)
    except Exception:
        return {}


def run_cuda(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    data, num, ang_max, cosmo, z_min = make_workload(args.workload)
    hps = sorted(data)

    # ---- cpu_baseline (N = 1, rank 0): the reference's Numba path in a fork pool, BEFORE this
    # process touches CUDA (a CUDA context does not survive a fork)
    cpu_frag = None
    if world == 1 and not args.no_cpu_baseline:
        cpu_frag = cpu_timed_steps(data, num, ang_max, 1, 1, args.cpu_tasks_per_core * 2,
                                   args.cpu_forests_per_task)

    import torch
    import torch.distributed as dist
    from picca_b200 import _corr, catalog, cf, dist as pdist, xcf
    from picca_b200.engine import MODE_AUTO, MODE_XCF, get_engine
    from picca_b200.params import params_from_module
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    os.environ["PICCA_B200_DEVICE"] = str(local_rank)
    eng = get_engine()

    configure(cf, data, num, ang_max)
    cf.userprint = xcf.userprint = lambda *a, **k: None   # ONE line on stdout: the JSON
    t0 = time.perf_counter()
    host = catalog.cached_pack(data)
    pack_s = time.perf_counter() - t0
    params = params_from_module(cf)
    nb = params.num_bins_r_par * params.num_bins_r_trans
    shard = pdist.Shard(eng, host, host, ang_max, world, rank)
    my_hps = [hps[k] for k in shard.mine]
    n_rows = len(my_hps)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        res = None
        for _ in range(steps):
            res = fn()
        ev1.record()
        barrier()
        ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=eng.device)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), res

    # ---- device-resident arm
    dev = eng.device_catalog(host)
    one_step = lambda: pdist.xi_sharded(eng, dev, dev, params, shard, MODE_AUTO)
    for _ in range(args.warmup):
        out = one_step()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = eng.launch_count()
    eng.lib.pb2_set_timing(1)
    kernel_ms = []

    def step_resident():
        o = one_step()
        kernel_ms.append(eng.lib.pb2_last_kernel_ms())
        return o
    total_ms, out = timed(step_resident, args.steps)
    eng.lib.pb2_set_timing(0)
    launches = eng.launch_count() - launches0
    sampler.stop_flag = True
    out_host = out.cpu().numpy() if out is not None else None   # rank 0: every row

    # ---- end-to-end arm: the plugin API on host data.  Every step drops the packed host and
    # device catalogues, so cf.fill_neighs packs the data dict again (host), uploads it (H2D),
    # builds the record copies and searches the neighbours; cf.compute_xi_batch runs the kernel
    # and copies the blocks back (D2H).  N > 1: blocks stay on the device for the gather.
    h2d = int(host.nbytes()) + shard.f1_index.nbytes + shard.rows.nbytes
    e2e_detail = {"api": "cf.fill_neighs + cf.compute_xi_batch on the data dict (pack + H2D + "
                         "neighbours + kernel + D2H per step)" if world == 1 else
                         "dist.BandShard + dist.xi_banded on the data dict (per rank: pack and H2D "
                         "of its band of HEALPix rows + halo, neighbours, kernel, NCCL gather, D2H)",
                  "pack_s_first": pack_s}

    def step_e2e():
        catalog.invalidate(data)
        eng.drop_catalogs()
        t_0 = time.perf_counter()
        if world == 1:
            catalog.cached_pack(data)
            e2e_detail["pack_s"] = time.perf_counter() - t_0
            cf.fill_neighs(my_hps)
            return cf.compute_xi_batch(my_hps)
        # N > 1: every rank packs and uploads only its band of HEALPix rows + the halo its
        # neighbour searches reach into (dist.BandShard), not the whole catalogue
        band = pdist.BandShard(eng, data, ang_max, world, rank)
        e2e_detail["pack_s"] = time.perf_counter() - t_0
        e2e_detail["band_rows"], e2e_detail["halo_rows"] = band.b1 - band.b0, band.h1 - band.h0
        e2e_detail["h2d_bytes_rank0"] = band.h2d_bytes
        t_1 = time.perf_counter()
        full = pdist.xi_banded(eng, band, params, MODE_AUTO)
        if full is None:
            return None
        if "pinned" not in e2e_detail:   # page-locked result buffer, allocated once
            e2e_detail["pinned"] = torch.empty(full.shape, dtype=full.dtype).pin_memory()
        e2e_detail["pinned"].copy_(full)
        e2e_detail["xi_gather_d2h_s"] = time.perf_counter() - t_1
        return e2e_detail["pinned"].numpy()
    if args.no_e2e:
        e2e_ms = total_ms
    else:
        dev = None
        step_e2e()
        e2e_ms, e2e_out = timed(step_e2e, args.e2e_steps)
        e2e_ms *= args.steps / args.e2e_steps
        if rank == 0:
            assert np.array_equal(e2e_out[:, 5].view(np.int64), out_host[:, 5].view(np.int64))
        host = catalog.cached_pack(data)
        dev = eng.device_catalog(host)

    # ---- parity of the timed result (rank 0, every N): whole rows against the oracle port
    parity = None
    if rank == 0 and not args.no_parity:
        parity = parity_check(data, num, ang_max, hps, out_host, args.parity_pixels,
                              os.cpu_count() or 1)

    # ---- distortion-matrix leg (BASELINE config 4: --rej 0.99 on the same sample, one reference
    # chunk seeded with the first HEALPix pixel as picca_dmat.py --nproc 1 does, :36,:471-485)
    dmat_info = None
    if not args.no_dmat:
        cf.reject = DMAT_REJECT
        cf.num_model_bins_r_par = CF_CFG["num_bins_r_par"]      # picca_dmat.py:114-122: coef 1
        cf.num_model_bins_r_trans = CF_CFG["num_bins_r_trans"]
        dparams = params_from_module(cf)
        dm_counts = {}

        def dmat_step(dev_cat):
            res, npall, npused = pdist.dmat_chunk_sharded(
                eng, dev_cat, dev_cat, dparams, shard, MODE_AUTO, DMAT_REJECT, hps[0],
                segments=args.dmat_segments)
            dm_counts["npall"], dm_counts["npused"] = npall, npused
            return res

        dmat_step(dev)
        dl0 = eng.launch_count()
        dm_ms, dm_res = timed(lambda: dmat_step(dev), args.dmat_steps)
        dm_launches = eng.launch_count() - dl0
        # one more, instrumented step outside the timed region: per-launch kernel times (the
        # event wait after every launch would serialise the host draw with the kernels) and the
        # in-kernel work counters
        eng.lib.pb2_set_timing(1)
        eng.collect_dmat_stats = True
        dm_kernel_ms = []
        eng.dmat_kernel_ms_log = dm_kernel_ms
        eng.sum_dmat_stats = None
        dmat_step(dev)
        eng.lib.pb2_set_timing(0)
        eng.collect_dmat_stats = False
        eng.dmat_kernel_ms_log = None
        dm_stats = dict(eng.sum_dmat_stats or {})
        eng.sum_dmat_stats = None
        dm_kms = float(np.sum(dm_kernel_ms))   # kernels of one step, this rank
        dm_nbytes = int(sum(t.numel() * 8 for t in dm_res))
        dm_host = [torch.empty(t.shape, dtype=torch.float64).pin_memory() for t in dm_res]

        def dmat_e2e():
            if world == 1:
                fresh = eng.device_catalog(host, pin=False, cache=False)
                r_ = dmat_step(fresh)
            else:   # band shards: pack + upload of the band, counts all-gathered
                band = pdist.BandShard(eng, data, ang_max, world, rank)
                r_, _, _ = pdist.dmat_chunk_banded(eng, band, dparams, MODE_AUTO, DMAT_REJECT,
                                                   hps[0], segments=args.dmat_segments)
            if rank == 0:
                for h_, t_ in zip(dm_host, r_):
                    h_.copy_(t_, non_blocking=True)
            return r_
        dm_e2e_ms, _ = timed(dmat_e2e, args.dmat_steps)
        used = dm_counts["npused"]
        step_s = dm_ms / args.dmat_steps * 1e-3
        dm_peak = eng.fp64_peak(8192)[0]
        ops_step = dm_stats.get("as_written_ops", 0.)
        pix_step = dm_stats.get("in_range_pixel_pairs", 0.)
        if world > 1:
            tot = torch.tensor([pix_step], dtype=torch.float64, device=eng.device)
            dist.all_reduce(tot)
            pix_total = float(tot.item())
        else:
            pix_total = pix_step
        # executed FP64 work and DRAM traffic of the distortion-matrix kernels on this workload,
        # from an ncu capture of the same launch sequence (scripts/gpu_dmat_profile.sh); divided
        # by the kernel time measured live here
        prof = load_profile_json("r02_dmat_counters.json")
        dm_roof = None
        if prof.get("workload") == args.workload and world == 1:
            ex = float(prof["fp64_thread_ops_per_step"])
            dm_roof = {"bound": "fp64", "achieved": ex / (dm_kms * 1e-3) / 1e12,
                       "peak": dm_peak / 1e12, "unit": "Tops/s (executed FP64 thread "
                       "instructions, 1 DFMA = 1 op)", "frac": ex / (dm_kms * 1e-3) / dm_peak,
                       "traffic": prof.get("dram_bytes_per_step"), "kernel_ms": dm_kms,
                       "source": "profiles/r02_dmat_counters.json (ncu) / live kernel time"}
        dmat_info = {
            "metric": "used forest pairs/sec (distortion matrix, --rej %.2f)" % DMAT_REJECT,
            "value": used / step_s, "unit": "forest pairs/s",
            "steps": args.dmat_steps, "ms_per_step": dm_ms / args.dmat_steps,
            "kernel_ms": dm_kms, "gpu_launches": int(dm_launches),
            "NPALL": dm_counts["npall"], "NPUSED": used,
            "in_range_pixel_pairs_per_step": pix_total,
            "pixel_pairs_per_s": pix_total / step_s,
            "dmat_shape": [int(dm_res[1].shape[0]), int(dm_res[1].shape[1])],
            "sum_dmat": float(dm_res[1].sum().item()),
            "sum_weights_dmat": float(dm_res[0].sum().item()),
            "e2e": {"value": used / (dm_e2e_ms / args.dmat_steps * 1e-3),
                    "unit": "forest pairs/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": dm_nbytes},
            "roofline": dm_roof,
            # the FP64 work of the reference algorithm AS WRITTEN (N_sel (15 U + 4) +
            # 40 N_inrange per used forest pair, cf.py:623-887) per second of this rank's
            # kernels: a statement about the algebra (the kernels contract per data bin and
            # execute far fewer operations), NOT a roofline -- see "roofline" above
            "as_written_ops_per_s": ops_step / (dm_kms * 1e-3) if dm_kms > 0 else None,
            "parallelism": ("--rej stream drawn on the host in %d segments overlapped with the "
                            "kernels; kept forest pairs sharded by owning HEALPix row x%d; ONE "
                            "NCCL all-reduce(SUM) of a flat buffer (matrix + 5 vectors)"
                            % (args.dmat_segments, world)),
        }

    # ---- forest x quasar cross-correlation leg (BASELINE config 3 at its surface density on
    # this footprint: 5/3 quasars per forest, picca_xcf.py default binning)
    xcf_info = None
    if not args.no_xcf:
        (objs, z_min2), nq = make_quasars(args.workload, cosmo)
        from picca_b200 import synth
        x_ang_max = synth.compute_ang_max(cosmo, XCF_CFG["r_trans_max"], z_min, z_min2)
        from tests import helpers
        helpers.configure(xcf, data, num, x_ang_max, objs=objs, **XCF_CFG)
        host_o = catalog.cached_pack(objs, is_object=True)
        dev_o = eng.device_catalog(host_o)
        xparams = params_from_module(xcf, cross=True)
        xnb = xparams.num_bins_r_par * xparams.num_bins_r_trans
        x_step = lambda: pdist.xi_sharded(eng, dev, dev_o, xparams, shard, MODE_XCF,
                                          cross_obj=True)
        for _ in range(max(1, args.warmup)):
            x_step()
        eng.lib.pb2_set_timing(1)
        x_kms = []

        def x_resident():
            o = x_step()
            x_kms.append(eng.lib.pb2_last_kernel_ms())
            return o
        xl0 = eng.launch_count()
        x_ms, x_out = timed(x_resident, args.xcf_steps)
        x_launches = eng.launch_count() - xl0
        eng.lib.pb2_set_timing(0)

        def x_e2e():
            catalog.invalidate(data)
            catalog.invalidate(objs)
            eng.drop_catalogs()
            xcf.fill_neighs(my_hps)
            if world == 1:
                return xcf.compute_xi_batch(my_hps)
            block = xcf.compute_xi_batch(my_hps, to_host=False)
            full = pdist.gather_rows(block, shard.mine, len(hps))
            return full.cpu().numpy() if full is not None else None
        dev = dev_o = None
        x_e2e()
        x_e2e_ms, _ = timed(x_e2e, 1)
        x_pairs = torch.zeros(1, dtype=torch.int64, device=eng.device)
        if rank == 0:
            x_pairs[0] = x_out[:, 5, :].view(torch.int64).sum()
        if world > 1:
            dist.broadcast(x_pairs, src=0)
        x_pairs = int(x_pairs.item())
        if rank == 0:
            xk = float(np.mean(x_kms))
            x_my = x_pairs if world == 1 else int(
                x_out[torch.as_tensor(shard.mine, device=eng.device), 5, :].view(
                    torch.int64).sum().item())
            x_peak = eng.fp64_peak(8192)[0]
            x_ach = FLOPS_PER_PAIR_XCF * x_my / (xk * 1e-3)
            xprof = load_profile_json("r02_xcf_counters.json")
            xcf_info = {
                "metric": "binned forest-pixel x quasar pairs/sec (xcf cross-correlation)",
                "value": x_pairs / (x_ms / args.xcf_steps * 1e-3), "unit": "pairs/s",
                "steps": args.xcf_steps, "ms_per_step": x_ms / args.xcf_steps,
                "gpu_launches": int(x_launches),
                "config": {"workload": "%s forests x %d quasars (BASELINE config 3 surface "
                                       "densities on this footprint)" % (args.workload, nq),
                           "np": 100, "nt": 50, "rp_min": -200., "rp_max": 200., "rt_max": 200.,
                           "binned_pairs_per_step": x_pairs},
                "e2e": {"value": x_pairs / (x_e2e_ms * 1e-3), "unit": "pairs/s",
                        "h2d_bytes_per_step": h2d + int(host_o.nbytes()),
                        "d2h_bytes_per_step": int(len(hps) * 6 * xnb * 8),
                        "api": "xcf.fill_neighs + xcf.compute_xi_batch on the data / objs dicts"},
                "roofline": {"bound": "fp64", "achieved": x_ach / 1e12, "peak": x_peak / 1e12,
                             "unit": "Tops/s (1 DFMA = 1 op)", "frac": x_ach / x_peak,
                             "ops_per_pair": FLOPS_PER_PAIR_XCF, "kernel": "pb2_xi_cross_chunk_t",
                             "kernel_ms": xk,
                             "traffic": xprof.get("dram_bytes_per_launch")
                             if xprof.get("workload") == args.workload and world == 1 else None},
            }

    # ---- totals (all ranks processed the whole sample between them)
    if rank == 0:
        pairs = int(out_host[:, 5].view(np.int64).sum())
        ms_step = total_ms / args.steps
        value = pairs / (ms_step * 1e-3)
        e2e_value = pairs / (e2e_ms / args.steps * 1e-3)
        peak_ops, _ = eng.fp64_peak(8192)
        kms = float(np.mean(kernel_ms)) if kernel_ms else float("nan")
        # rank 0's kernel handles its own shard: scale pairs by its share of the work
        my_pairs = int(out_host[shard.mine, 5].view(np.int64).sum())
        achieved = FLOPS_PER_PAIR * my_pairs / (kms * 1e-3)
        peaks = load_profile_json(os.path.join("..", "MEASURED_PEAKS.json"))
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        # DRAM bytes of one launch of the dominant kernel on this workload, from an ncu capture
        # (scripts/gpu_traffic.sh); null if no capture matches
        traffic = None
        for name in ("r02_traffic.json", "r01_traffic.json"):
            tr = load_profile_json(name)
            if tr.get("workload") == args.workload and world == 1:
                traffic = tr.get("dram_bytes_per_launch")
                break
        alg_bytes = 48.0 * host.n_pix + n_rows * 6 * nb * 8.0
        line = {
            "metric": "binned forest-pixel pairs/sec (cf auto-correlation)",
            "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "forests": host.n_los, "pixels": host.n_pix,
                       "healpix": len(hps), "np": 50, "nt": 50, "rp_max": 200., "rt_max": 200.,
                       "nside": 32, "binned_pairs_per_step": pairs,
                       "angles": "host (NumPy arccos, parity mode)" if _corr.HOST_ANGLES else
                       "device (acos/sin/cos of pb2_neigh.cu; PICCA_B200_HOST_ANGLES=1 is the "
                       "parity mode)",
                       "l2_policy": "inputs (%.2f GB in HBM) larger than L2"
                                    % (eng.device_catalog(host).device_bytes() / 1e9),
                       "parallelism": "healpix LPT shards x%d, gather to rank 0" % world},
            "e2e": {"value": e2e_value, "unit": "pairs/s",
                    "h2d_bytes_per_step": e2e_detail.get("h2d_bytes_rank0", h2d),
                    "d2h_bytes_per_step": int(n_rows * 6 * nb * 8), "steps": args.e2e_steps,
                    **{k: v for k, v in e2e_detail.items() if k != "pinned"}},
            "gpu_launches": int(launches),
            "clocks": sampler.summary(),
            "roofline": {"bound": "fp64", "achieved": achieved / 1e12, "peak": peak_ops / 1e12,
                         "unit": "Tops/s (1 DFMA = 1 op)", "frac": achieved / peak_ops,
                         "traffic": traffic, "kernel": "pb2_xi_auto_diag", "kernel_ms": kms,
                         "peak_source": "pb2_fp64_peak DFMA microbenchmark, measured in this run "
                                        "(MEASURED_PEAKS.json has no fp64 entry)",
                         "ops_per_pair": FLOPS_PER_PAIR},
            "roofline_hbm": {"bound": "hbm", "achieved": alg_bytes / (kms * 1e-3) / 1e9,
                             "peak": hbm_peak, "unit": "GB/s",
                             "frac": alg_bytes / (kms * 1e-3) / 1e9 / hbm_peak,
                             "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"},
        }
        if parity is not None:
            line["parity_check"] = parity
        if cpu_frag is not None:
            line["cpu_baseline"] = {k: cpu_frag[k] for k in ("value", "unit", "cores", "kind",
                                                             "sample")}
        if dmat_info is not None:
            line["dmat"] = dmat_info
        if xcf_info is not None:
            line["xcf"] = xcf_info
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_banded(args):
    """DESI-scale arm (BASELINE configs[4], ``--workload c5_1m``): auto-correlation + distortion
    matrix with BAND shards only.  A rank draws the light index of the whole survey (positions and
    shapes, vectorised), then generates, packs and uploads just its band of HEALPix rows + halo
    (``synth.make_forest_band`` behind ``dist.BandShard(index=..., band_source=...)``): no rank
    ever holds the 1M-forest catalogue.  Same JSON contract as the default arm; ``e2e`` = pack +
    H2D of the band + neighbours + kernel + NCCL gather + D2H from the band's data dict."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    from picca_b200 import synth
    kw = dict(WORKLOADS[args.workload])
    n_forest = kw.pop("n_forest")
    t_wall0 = time.perf_counter()
    ix = synth.make_forest_index(n_forest, **kw)
    ang_max = synth.compute_ang_max(ix.cosmo, CF_CFG["r_trans_max"], ix.z_min)

    import torch
    import torch.distributed as dist
    from picca_b200 import _corr, catalog, cf, dist as pdist
    from picca_b200.engine import MODE_AUTO, get_engine
    from picca_b200.params import params_from_module
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    os.environ["PICCA_B200_DEVICE"] = str(local_rank)
    eng = get_engine()
    idx = pdist.RowIndex.from_arrays(ix.healpixs, ix.counts, ix.xyz, ix.npix)
    hps = ix.healpixs
    threads = max(1, (os.cpu_count() or 1) // max(world, 1))
    t0 = time.perf_counter()
    band = pdist.BandShard(eng, None, ang_max, world, rank, index=idx,
                           band_source=lambda h0, h1: synth.make_forest_band(ix, h0, h1, threads=threads))
    setup_s = time.perf_counter() - t0
    data = band.data
    configure(cf, data, n_forest, ang_max)
    cf.userprint = lambda *a, **k: None
    params = params_from_module(cf)
    nb = params.num_bins_r_par * params.num_bins_r_trans

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        res = None
        for _ in range(steps):
            res = fn()
        ev1.record()
        barrier()
        ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=eng.device)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), res

    one_step = lambda: pdist.xi_banded(eng, band, params, MODE_AUTO)
    for _ in range(args.warmup):
        one_step()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = eng.launch_count()
    eng.lib.pb2_set_timing(1)
    kernel_ms = []

    def step():
        o = one_step()
        kernel_ms.append(eng.lib.pb2_last_kernel_ms())
        return o
    total_ms, out = timed(step, args.steps)
    eng.lib.pb2_set_timing(0)
    launches = eng.launch_count() - launches0
    sampler.stop_flag = True
    out_host = out.cpu().numpy() if out is not None else None

    # ---- end to end from the band's data dict: pack + H2D + neighbours + kernel + gather + D2H
    e2e_detail = {"api": "dist.BandShard + dist.xi_banded on the band's data dict (per rank: pack "
                         "and H2D of its band of HEALPix rows + halo, neighbours, kernel, NCCL "
                         "gather, D2H); generating the synthetic band is not included"}
    if args.no_e2e:
        e2e_ms = total_ms
    else:
        band_data = band.data
        band = None
        eng.drop_catalogs()

        def step_e2e():
            t_0 = time.perf_counter()
            b = pdist.BandShard(eng, None, ang_max, world, rank, index=idx,
                                band_source=lambda h0, h1: band_data)
            e2e_detail["pack_s"] = time.perf_counter() - t_0
            e2e_detail["band_rows"], e2e_detail["halo_rows"] = b.b1 - b.b0, b.h1 - b.h0
            e2e_detail["h2d_bytes_rank0"] = b.h2d_bytes
            full = pdist.xi_banded(eng, b, params, MODE_AUTO)
            step_e2e.band = b
            return full.cpu().numpy() if full is not None else None
        e2e_ms, e2e_out = timed(step_e2e, 1)
        e2e_ms *= args.steps
        band = step_e2e.band
        if rank == 0:
            assert np.array_equal(e2e_out[:, 5].view(np.int64), out_host[:, 5].view(np.int64))

    # ---- parity (rank 0): whole rows from the middle of its band against the oracle port; the
    # oracle sees the rows within reach of the chosen ones (all inside this rank's halo)
    parity = None
    if rank == 0 and not args.no_parity:
        mid = (band.b0 + band.b1) // 2
        chosen_rows = np.arange(mid, mid + max(1, args.parity_pixels))
        near = np.nonzero(idx.near(idx, chosen_rows, ang_max).any(axis=0))[0]
        assert near.min() >= band.h0 and near.max() < band.h1
        sub = {hps[r]: data[hps[r]] for r in near}
        sub_hps = sorted(sub)
        rows_sub = np.zeros((len(sub_hps), 6, nb))
        for r in chosen_rows:
            rows_sub[sub_hps.index(hps[r])] = out_host[r]
        parity = parity_check(sub, n_forest, ang_max, sub_hps, rows_sub, len(chosen_rows),
                              os.cpu_count() or 1, chosen=[hps[r] for r in chosen_rows])

    # ---- distortion matrix, --rej 0.99, one reference chunk seeded with the first pixel
    dmat_info = None
    if not args.no_dmat:
        cf.reject = DMAT_REJECT
        cf.num_model_bins_r_par = CF_CFG["num_bins_r_par"]
        cf.num_model_bins_r_trans = CF_CFG["num_bins_r_trans"]
        dparams = params_from_module(cf)
        counts = {}

        def dmat_step():
            res, npall, npused = pdist.dmat_chunk_banded(eng, band, dparams, MODE_AUTO, DMAT_REJECT,
                                                         hps[0], segments=args.dmat_segments)
            counts["npall"], counts["npused"] = npall, npused
            return res
        if args.warmup:
            dmat_step()
        dl0 = eng.launch_count()
        dm_ms, dm_res = timed(dmat_step, args.dmat_steps)
        dmat_info = {
            "metric": "used forest pairs/sec (distortion matrix, --rej %.2f)" % DMAT_REJECT,
            "value": counts["npused"] / (dm_ms / args.dmat_steps * 1e-3), "unit": "forest pairs/s",
            "steps": args.dmat_steps, "ms_per_step": dm_ms / args.dmat_steps,
            "gpu_launches": int(eng.launch_count() - dl0), "NPALL": counts["npall"],
            "NPUSED": counts["npused"],
            "dmat_shape": [int(dm_res[1].shape[0]), int(dm_res[1].shape[1])],
            "sum_dmat": float(dm_res[1].sum().item()),
            "sum_weights_dmat": float(dm_res[0].sum().item()),
            "parallelism": "band shards x%d: neighbour counts all-gathered, --rej stream drawn on "
                           "the host in %d segments overlapped with the kernels, ONE NCCL "
                           "all-reduce(SUM) of a flat buffer" % (world, args.dmat_segments)}

    peak_ops, _ = eng.fp64_peak(8192)
    kms_all = torch.zeros(world, dtype=torch.float64, device=eng.device)
    kms_all[rank] = float(np.mean(kernel_ms))
    if world > 1:
        dist.all_reduce(kms_all)
    if rank == 0:
        pairs = int(out_host[:, 5].view(np.int64).sum())
        ms_step = total_ms / args.steps
        kms = float(np.mean(kernel_ms))
        my_pairs = int(out_host[band.b0:band.b1, 5].view(np.int64).sum())
        achieved = FLOPS_PER_PAIR * my_pairs / (kms * 1e-3)
        line = {
            "metric": "binned forest-pixel pairs/sec (cf auto-correlation)",
            "value": pairs / (ms_step * 1e-3), "unit": "pairs/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": args.workload, "forests": int(n_forest),
                       "pixels": int(ix.npix.sum()), "healpix": len(hps), "np": 50, "nt": 50,
                       "rp_max": 200., "rt_max": 200., "nside": 32, "binned_pairs_per_step": pairs,
                       "angles": "host (NumPy arccos, parity mode)" if _corr.HOST_ANGLES else
                       "device (acos/sin/cos of pb2_neigh.cu)",
                       "l2_policy": "inputs (%.2f GB in HBM on rank 0) larger than L2"
                                    % (band.dev.device_bytes() / 1e9),
                       "parallelism": "band shards x%d (contiguous HEALPix rows + halo per rank), "
                                      "gather to rank 0" % world,
                       "rank0_forests_held": int(band.host.n_los),
                       "kernel_ms_per_rank": [round(float(x), 1) for x in kms_all.tolist()],
                       "generate_pack_upload_s_rank0": setup_s,
                       "wall_s_total": time.perf_counter() - t_wall0},
            "e2e": ({"value": None, "unit": "pairs/s", "note": "--no-e2e: not measured in this run; "
                     "config.generate_pack_upload_s_rank0 and config.wall_s_total bound it",
                     "h2d_bytes_per_step": int(band.h2d_bytes),
                     "d2h_bytes_per_step": int(len(hps) * 6 * nb * 8)} if args.no_e2e else
                    {"value": pairs / (e2e_ms / args.steps * 1e-3), "unit": "pairs/s",
                     "h2d_bytes_per_step": e2e_detail.get("h2d_bytes_rank0"),
                     "d2h_bytes_per_step": int(len(hps) * 6 * nb * 8), "steps": 1, **e2e_detail}),
            "gpu_launches": int(launches), "clocks": sampler.summary(),
            "roofline": {"bound": "fp64", "achieved": achieved / 1e12, "peak": peak_ops / 1e12,
                         "unit": "Tops/s (1 DFMA = 1 op)", "frac": achieved / peak_ops,
                         "traffic": None, "kernel": "pb2_xi_auto_diag", "kernel_ms": kms,
                         "peak_source": "pb2_fp64_peak DFMA microbenchmark, measured in this run",
                         "ops_per_pair": FLOPS_PER_PAIR},
        }
        if parity is not None:
            line["parity_check"] = parity
        if dmat_info is not None:
            line["dmat"] = dmat_info
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--workload", default="c2_100k", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-tasks-per-core", type=int, default=4)
    ap.add_argument("--cpu-forests-per-task", type=int, default=2)
    ap.add_argument("--parity-pixels", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-dmat", action="store_true", help="skip the distortion-matrix leg")
    ap.add_argument("--no-xcf", action="store_true", help="skip the cross-correlation leg")
    ap.add_argument("--no-e2e", action="store_true",
                    help="probe runs only: skip the host-buffer arm (e2e repeats `value`)")
    ap.add_argument("--e2e-steps", type=int, default=0,
                    help="timed steps of the end-to-end arm (default: min(steps, 3); its time is "
                         "scaled to --steps)")
    ap.add_argument("--dmat-steps", type=int, default=2)
    ap.add_argument("--dmat-segments", type=int, default=8)
    ap.add_argument("--xcf-steps", type=int, default=3)
    ap.add_argument("--banded", action="store_true",
                    help="band shards only: every rank generates / holds just its band of HEALPix "
                         "rows + halo (always on for --workload c5_1m)")
    args = ap.parse_args()
    if args.e2e_steps <= 0:
        args.e2e_steps = min(args.steps, 3)
    if args.impl == "reference":
        run_reference(args)
    elif args.banded or args.workload == "c5_1m":
        run_banded(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
