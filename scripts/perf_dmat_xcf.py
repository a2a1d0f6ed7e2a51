"""GPU probe: distortion matrix (--rej 0.99, BASELINE config 4 shape) and forest x quasar
cross-correlation (config 3 shape) on reduced footprints of the same surface density."""
import argparse
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from picca_b200 import cf, xcf, synth  # noqa: E402
from picca_b200.engine import get_engine  # noqa: E402
from tests import helpers  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=5000)
ap.add_argument("--nq", type=int, default=8300)
ap.add_argument("--side", type=float, default=15.5)
ap.add_argument("--rej", type=float, default=0.99)
args = ap.parse_args()

data, num, z_min, z_max, cosmo = synth.make_forests(
    args.n, seed=20260102, nside=32, ra_deg=(0., args.side), dec_deg=(0., args.side))
eng = get_engine()
torch = eng.torch
eng.lib.pb2_set_timing(1)
hps = sorted(data)

# ---- dmat
ang_max = synth.compute_ang_max(cosmo, 200., z_min)
helpers.configure(cf, data, num, ang_max, num_bins_r_par=50, num_bins_r_trans=50, r_par_max=200.,
                  r_trans_max=200., num_model_bins_r_par=50, num_model_bins_r_trans=50, nside=32,
                  reject=args.rej)
for rep in range(2):
    cf.fill_neighs(hps)
    np.random.seed(hps[0])
    torch.cuda.synchronize()
    t0 = time.time()
    res = cf.compute_dmat(hps)
    torch.cuda.synchronize()
    dt = time.time() - t0
    print("dmat rep %d: %d forests, NPALL %d NPUSED %d, %.3fs wall, kernel %.1f ms -> %.1f used "
          "forest pairs/s, sum(dmat)=%.6e" % (rep, num, res[6], res[7], dt,
                                              eng.lib.pb2_last_kernel_ms(),
                                              res[7] / (eng.lib.pb2_last_kernel_ms() * 1e-3),
                                              res[1].sum()))

# ---- xcf
objs, z_min2 = synth.make_quasars(args.nq, seed=20260103, nside=32, ra_deg=(0., args.side),
                                  dec_deg=(0., args.side), cosmo=cosmo)
ang_max = synth.compute_ang_max(cosmo, 200., z_min, z_min2)
helpers.configure(xcf, data, num, ang_max, objs=objs, num_bins_r_par=100, num_bins_r_trans=50,
                  r_par_max=200., r_par_min=-200., r_trans_max=200., nside=32, alpha_obj=1.44,
                  num_model_bins_r_par=100, num_model_bins_r_trans=50, reject=args.rej)
for rep in range(2):
    torch.cuda.synchronize()
    t0 = time.time()
    xcf.fill_neighs(hps)
    out = xcf.compute_xi_batch(hps, normalise=False)
    torch.cuda.synchronize()
    dt = time.time() - t0
    pairs = int(out[:, 5, :].view(np.int64).sum())
    kms = eng.lib.pb2_last_kernel_ms()
    print("xcf rep %d: %d forests x %d quasars, %d binned pairs, %.3fs wall, kernel %.1f ms -> "
          "%.3e pairs/s" % (rep, num, args.nq, pairs, dt, kms, pairs / (kms * 1e-3)))
for rep in range(2):
    xcf.fill_neighs(hps)
    np.random.seed(hps[0])
    torch.cuda.synchronize()
    t0 = time.time()
    res = xcf.compute_dmat(hps)
    torch.cuda.synchronize()
    dt = time.time() - t0
    print("xdmat rep %d: NPALL %d NPUSED %d, %.3fs wall, kernel %.1f ms" % (
        rep, res[6], res[7], dt, eng.lib.pb2_last_kernel_ms()))
