#!/usr/bin/env python
"""Benchmark of the forest pair-counting hot path (BASELINE.json metric: binned forest-pixel
pairs per second, auto-correlation, synthetic 100k-forest DR16-like sample, nside 32).

    python bench.py --gpus N --steps K --warmup W            # CUDA path (picca_b200)
    python bench.py --impl reference --steps K --warmup W    # CPU arm: the oracle port, all cores

A step is one pass of the hot path over the workload: device neighbour search (fill_neighs) +
pair kernel (compute_xi) + normalisation, per-HEALPix blocks produced for every pixel.  `value`
is timed with the packed catalogue already resident in HBM; `e2e` repeats the steps through the
host-buffer entry (pinned host catalogue -> H2D -> record packing + kernels -> D2H of the blocks),
copies inside the timed region.  N > 1: HEALPix pixels are LPT-partitioned over the ranks (strong scaling),
blocks gathered to rank 0 inside the timed region, time = max over ranks.
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
}
CF_CFG = dict(num_bins_r_par=50, num_bins_r_trans=50, r_par_max=200., r_par_min=0.,
              r_trans_max=200., nside=32)
DMAT_REJECT = 0.99  # BASELINE.json configs[3]
FLOPS_PER_PAIR = 30.  # algorithmic FP64 ops per binned pair (cf)


def make_workload(name):
    from picca_b200 import synth
    kw = dict(WORKLOADS[name])
    n = kw.pop("n_forest")
    data, num, z_min, z_max, cosmo = synth.make_forests(n, **kw)
    ang_max = synth.compute_ang_max(cosmo, CF_CFG["r_trans_max"], z_min)
    return data, num, ang_max


class Cfg:
    """stand-in for the module globals of picca.cf (what picca_cf.py assigns, :343-368)"""


def configure(mod, data, num, ang_max):
    from tests import helpers
    helpers.configure(mod, data, num, ang_max, **CF_CFG)


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
# CPU arm / cpu_baseline: the oracle port (C, pthreads) on a bounded sample of the same workload
# --------------------------------------------------------------------------------------------
def cpu_sample_run(data, num, ang_max, n_healpix, threads):
    """Times the oracle's compute_xi loop (C batch driver, all host threads) on `n_healpix`
    HEALPix pixels of the workload.  Returns (binned pairs, seconds, description)."""
    import ctypes
    from oracle import _host, _kernels
    from picca_b200 import catalog
    cfg = Cfg()
    configure(cfg, data, num, ang_max)
    host = catalog.cached_pack(data)
    A = host.arrays
    hps = host.healpixs
    # centre of the footprint: pixels with typical neighbour counts
    mid = len(hps) // 2
    chosen = hps[mid:mid + n_healpix]
    cat = _host.catalogue(data)
    f1_index, rows, nb_off, nb_idx, nb_ang = [], [], [0], [], []
    for r, hp in enumerate(chosen):
        a, b = host.first_of(hp)
        for f1 in range(a, b):
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
    out = np.zeros((len(chosen), 6, nb))
    f1_index = np.array(f1_index, dtype=np.int64)
    rows = np.array(rows, dtype=np.int64)
    nb_off = np.array(nb_off, dtype=np.int64)
    nb_idx = np.concatenate(nb_idx).astype(np.int64)
    nb_ang = np.concatenate(nb_ang).astype(np.float64)
    lib = _kernels.lib()
    dp, lp = _kernels.dp, _kernels.lp
    delta = np.where(A["weights"] != 0, A["delta_w"] / np.where(A["weights"] != 0, A["weights"], 1.),
                     0.)
    t0 = time.perf_counter()
    lib.orc_xi_auto_batch(
        ctypes.byref(p), lp(A["offset"]), dp(A["z"]), dp(A["r_comov"]), dp(A["dist_m"]),
        dp(A["weights"]), dp(delta), dp(A["z_qso"]), lp(A["offset"]), dp(A["z"]),
        dp(A["r_comov"]), dp(A["dist_m"]), dp(A["weights"]), dp(delta), dp(A["z_qso"]),
        ctypes.c_int64(len(f1_index)), lp(f1_index), lp(rows), lp(nb_off), lp(nb_idx),
        dp(nb_ang), None, ctypes.c_int64(len(chosen)), dp(out), ctypes.c_int32(threads))
    dt = time.perf_counter() - t0
    pairs = int(out[:, 5, :].view(np.int64).sum())
    desc = "%d of %d HEALPix pixels (%d forests, %d forest pairs) of the workload" % (
        len(chosen), len(hps), len(f1_index), nb_idx.size)
    return pairs, dt, desc


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port; the reference itself is
    Python+Numba and cannot travel to the GPU box) with every host thread, bounded sample/step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    data, num, ang_max = make_workload(args.workload)
    n_hp = args.cpu_healpix
    times, pairs, desc = [], 0, ""
    for step in range(args.warmup + args.steps):
        pairs, dt, desc = cpu_sample_run(data, num, ang_max, n_hp, threads)
        if step >= args.warmup:
            times.append(dt)
    ms = 1e3 * float(np.mean(times))
    value = pairs / (ms * 1e-3)
    line = {
        "impl": "reference", "metric": "binned forest-pixel pairs/sec (cf auto-correlation)",
        "value": value, "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "np": 50, "nt": 50, "rp_max": 200., "rt_max": 200.,
                   "nside": 32},
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": threads, "kind": "port",
                         "sample": desc},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------
# CUDA arm
# --------------------------------------------------------------------------------------------
def run_cuda(args):
    import torch
    import torch.distributed as dist
    from picca_b200 import catalog, cf, dist as pdist
    from picca_b200.engine import MODE_AUTO, get_engine
    from picca_b200.params import params_from_module

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    os.environ["PICCA_B200_DEVICE"] = str(local_rank)
    eng = get_engine()

    data, num, ang_max = make_workload(args.workload)
    configure(cf, data, num, ang_max)
    host = catalog.cached_pack(data)
    params = params_from_module(cf)
    nb = params.num_bins_r_par * params.num_bins_r_trans
    hps = host.healpixs

    # ---- shard: LPT over estimated pair work (identical on every rank)
    work = pdist.estimate_work(host, host, ang_max)
    mine = pdist.lpt_partition(work, world)[rank]
    my_hps = [hps[k] for k in mine]
    f1_parts = [np.arange(*host.first_of(hp), dtype=np.int32) for hp in my_hps]
    f1_index = np.concatenate(f1_parts) if f1_parts else np.zeros(0, np.int32)
    rows = np.concatenate([np.full(len(p), k, np.int32) for k, p in enumerate(f1_parts)]) \
        if f1_parts else np.zeros(0, np.int32)
    d_f1 = torch.as_tensor(f1_index, device=eng.device)
    d_rows = torch.as_tensor(rows, device=eng.device)
    n_rows = len(my_hps)

    def one_step(dev_cat):
        pairs = eng.neighbours(dev_cat, dev_cat, params, MODE_AUTO, d_f1)
        out = eng.xi(dev_cat, dev_cat, params, pairs, d_rows, n_rows, normalise=True)
        if world > 1:
            out = pdist.gather_rows(out, mine, len(hps))
        return out

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
    for _ in range(args.warmup):
        out = one_step(dev)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = eng.launch_count()
    eng.lib.pb2_set_timing(1)
    kernel_ms = []

    def step_resident():
        out = one_step(dev)
        kernel_ms.append(eng.lib.pb2_last_kernel_ms())
        return out
    total_ms, out = timed(step_resident, args.steps)
    eng.lib.pb2_set_timing(0)
    launches = eng.launch_count() - launches0
    sampler.stop_flag = True

    # ---- end-to-end arm: pinned host catalogue -> H2D -> kernels -> D2H of the blocks
    pinned = {k: torch.from_numpy(v).pin_memory() for k, v in host.arrays.items()}
    h2d = int(sum(v.nbytes for v in host.arrays.values())) + f1_index.nbytes + rows.nbytes
    result_host = torch.empty((n_rows if world == 1 else len(hps), 6, nb), dtype=torch.float64).pin_memory()

    def step_e2e():
        fresh = catalog.DeviceCatalog.from_tensors(
            host, eng.device, {k: v.to(eng.device, non_blocking=True) for k, v in pinned.items()})
        o = one_step(fresh)
        if o is not None:
            result_host[:o.shape[0]].copy_(o, non_blocking=True)
        return o
    if args.no_e2e:
        e2e_ms = total_ms
    else:
        step_e2e()
        e2e_ms, _ = timed(step_e2e, args.steps)

    # ---- distortion-matrix leg (BASELINE config 4: --rej 0.99 on the same sample, one reference
    # chunk seeded with the first HEALPix pixel as picca_dmat.py --nproc 1 does, :36,:471-485)
    dmat_info = None
    if not args.no_dmat:
        all_f1 = torch.as_tensor(np.arange(host.n_los, dtype=np.int32), device=eng.device)
        row_of_f1 = np.repeat(np.arange(len(hps)), np.diff(host.arrays["hp_first"]))
        row_owner = np.zeros(len(hps), dtype=np.int64)
        for r_, part in enumerate(pdist.lpt_partition(work, world)):
            row_owner[part] = r_
        cf.reject = DMAT_REJECT
        cf.num_model_bins_r_par = CF_CFG["num_bins_r_par"]      # picca_dmat.py:114-122: coef 1
        cf.num_model_bins_r_trans = CF_CFG["num_bins_r_trans"]
        dparams = params_from_module(cf)
        dm_counts = {}

        def dmat_step(dev_cat):
            # neighbour search is replicated (ms); the draw is identical on every rank (Q6);
            # the kernels run on this rank's share; one NCCL all-reduce sums the accumulators
            prs = eng.neighbours(dev_cat, dev_cat, dparams, MODE_AUTO, all_f1)
            keep = pdist.draw_keep_mask(prs.n_pairs, DMAT_REJECT, hps[0])
            dm_counts["npall"], dm_counts["npused"] = int(prs.n_pairs), int(keep.sum())
            if world > 1:
                f1_of_pair = prs.nb_f1.cpu().numpy()
                keep = pdist.shard_keep_mask(keep, row_of_f1[f1_of_pair], row_owner, rank)
            dm_counts["pairs"], dm_counts["keep"] = prs, keep
            return pdist.dmat_sharded(eng, dev_cat, dev_cat, dparams, prs, keep, world=world)

        dmat_step(dev)
        eng.lib.pb2_set_timing(1)
        eng.collect_dmat_stats = True
        dm_kernel_ms = []

        def dmat_resident():
            r_ = dmat_step(dev)
            dm_kernel_ms.append(eng.lib.pb2_last_kernel_ms())
            return r_
        dl0 = eng.launch_count()
        dm_ms, dm_res = timed(dmat_resident, args.dmat_steps)
        dm_launches = eng.launch_count() - dl0
        eng.lib.pb2_set_timing(0)
        eng.collect_dmat_stats = False
        dm_stats = dict(eng.last_dmat_stats or {})
        dm_nbytes = int(sum(t.numel() * 8 for t in dm_res))
        dm_host = [torch.empty(t.shape, dtype=torch.float64).pin_memory() for t in dm_res]

        def dmat_e2e():
            fresh = catalog.DeviceCatalog.from_tensors(
                host, eng.device,
                {k: v.to(eng.device, non_blocking=True) for k, v in pinned.items()})
            r_ = dmat_step(fresh)
            if rank == 0:
                for h_, t_ in zip(dm_host, r_):
                    h_.copy_(t_, non_blocking=True)
            return r_
        dm_e2e_ms, _ = timed(dmat_e2e, args.dmat_steps)
        used = dm_counts["npused"]
        # pixel pairs the matrix was built from: the binned pairs of this rank's kept forest pairs
        # (one launch of the xi kernel over that sub-list, outside the timed region)
        sub = dm_counts["pairs"].subset(dm_counts["keep"])
        cnt = eng.xi(dev, dev, dparams, sub, torch.zeros(sub.n_f1, dtype=torch.int32,
                                                         device=eng.device), 1)
        dm_pix = cnt[:, 5, :].view(torch.int64).sum().reshape(1).clone()
        if world > 1:
            dist.all_reduce(dm_pix)
        dm_pix = int(dm_pix.item())
        dm_peak = eng.fp64_peak(8192)[0]
        dm_ops_s = dm_stats.get("as_written_ops", 0.) / (float(np.mean(dm_kernel_ms)) * 1e-3)
        dm_as_written = {"ops_per_step_this_rank": dm_stats.get("as_written_ops"),
                         "ops_per_s": dm_ops_s, "fp64_peak_ops_per_s": dm_peak,
                         "frac_of_fp64_peak": dm_ops_s / dm_peak,
                         "unique_model_bins_per_used_pair":
                             dm_stats.get("sum_unique_model_bins", 0.) / max(1, used // world)}
        dmat_info = {
            "metric": "used forest pairs/sec (distortion matrix, --rej %.2f)" % DMAT_REJECT,
            "value": used / (dm_ms / args.dmat_steps * 1e-3), "unit": "forest pairs/s",
            "steps": args.dmat_steps, "ms_per_step": dm_ms / args.dmat_steps,
            "kernel_ms": float(np.mean(dm_kernel_ms)), "gpu_launches": int(dm_launches),
            "NPALL": dm_counts["npall"], "NPUSED": used, "pixel_pairs_per_step": dm_pix,
            "pixel_pairs_per_s": dm_pix / (dm_ms / args.dmat_steps * 1e-3),
            "dmat_shape": [int(dm_res[1].shape[0]), int(dm_res[1].shape[1])],
            "sum_dmat": float(dm_res[1].sum().item()),
            "sum_weights_dmat": float(dm_res[0].sum().item()),
            "e2e": {"value": used / (dm_e2e_ms / args.dmat_steps * 1e-3),
                    "unit": "forest pairs/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": dm_nbytes},
            # the FP64 work of the reference algorithm AS WRITTEN (N_sel (15 U + 4) +
            # 40 N_inrange per used forest pair, cf.py:623-887) over this rank's kernel time; the
            # kernels contract per data bin instead and execute far fewer
            # operations, so this can exceed the machine's peak: it measures the algebra, the
            # used-forest-pair rate above measures the kernel
            "as_written": dm_as_written,
            "parallelism": "kept forest pairs sharded by owning HEALPix row x%d, "
                           "NCCL all-reduce(SUM) of dmat + 5 vectors" % world if world > 1 else
                           "one GPU, one reference chunk",
        }

    # ---- totals (all ranks processed the whole sample between them)
    local_pairs = torch.tensor([0], dtype=torch.int64, device=eng.device)
    if world == 1:
        local_pairs[0] = out[:, 5, :].view(torch.int64).sum()
    elif rank == 0:
        local_pairs[0] = out[:, 5, :].view(torch.int64).sum()
    if world > 1:
        dist.broadcast(local_pairs, src=0)
    pairs = int(local_pairs.item())

    if rank == 0:
        ms_step = total_ms / args.steps
        value = pairs / (ms_step * 1e-3)
        e2e_value = pairs / (e2e_ms / args.steps * 1e-3)
        peak_ops, _ = eng.fp64_peak(8192)
        kms = float(np.mean(kernel_ms)) if kernel_ms else float("nan")
        # rank 0's kernel handles its own shard: scale pairs by its share of the work
        my_pairs = pairs if world == 1 else int(
            out[torch.as_tensor(mine, device=eng.device), 5, :].view(torch.int64).sum().item())
        achieved = FLOPS_PER_PAIR * my_pairs / (kms * 1e-3)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        # DRAM bytes of one launch of the dominant kernel on this workload, from an ncu capture
        # (scripts/gpu_traffic.sh -> profiles/r01_traffic.json); null if no capture matches
        traffic = None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
            if tr.get("workload") == args.workload and world == 1:
                traffic = tr.get("dram_bytes_per_launch")
        except Exception:
            pass
        alg_bytes = 48.0 * host.n_pix + n_rows * 6 * nb * 8.0
        line = {
            "metric": "binned forest-pixel pairs/sec (cf auto-correlation)",
            "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "forests": host.n_los, "pixels": host.n_pix,
                       "healpix": len(hps), "np": 50, "nt": 50, "rp_max": 200., "rt_max": 200.,
                       "nside": 32, "binned_pairs_per_step": pairs,
                       "l2_policy": "inputs (%.2f GB in HBM) larger than L2" % (dev.device_bytes() / 1e9),
                       "parallelism": "healpix LPT shards x%d, gather to rank 0" % world},
            "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": int(result_host.numel() * 8)},
            "gpu_launches": int(launches),
            "clocks": sampler.summary(),
            "roofline": {"bound": "fp64", "achieved": achieved / 1e12, "peak": peak_ops / 1e12,
                         "unit": "Tops/s (1 DFMA = 1 op)", "frac": achieved / peak_ops,
                         "traffic": traffic, "kernel": "pb2_xi_auto_diag", "kernel_ms": kms,
                         "peak_source": "pb2_fp64_peak DFMA microbenchmark, measured in this run",
                         "ops_per_pair": FLOPS_PER_PAIR},
            "roofline_hbm": {"bound": "hbm", "achieved": alg_bytes / (kms * 1e-3) / 1e9,
                             "peak": hbm_peak, "unit": "GB/s",
                             "frac": alg_bytes / (kms * 1e-3) / 1e9 / hbm_peak,
                             "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"},
        }
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            cpairs, cdt, cdesc = cpu_sample_run(data, num, ang_max, args.cpu_healpix, threads)
            line["cpu_baseline"] = {"value": cpairs / cdt, "unit": "pairs/s", "cores": threads,
                                    "kind": "port", "sample": cdesc}
        if dmat_info is not None:
            line["dmat"] = dmat_info
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--workload", default="c2_100k", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-healpix", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-dmat", action="store_true", help="skip the distortion-matrix leg")
    ap.add_argument("--no-e2e", action="store_true",
                    help="probe runs only: skip the host-buffer arm (e2e repeats `value`)")
    ap.add_argument("--dmat-steps", type=int, default=2)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
