"""BASELINE config 3 at one third of its size and the same surface densities on one GPU:
forest x quasar cross-correlation, 100k forests x 170k quasars on the 4760 deg^2 footprint of the
100k-forest sample (config 3: 300k x 500k on ~14 000 deg^2), picca_xcf.py default binning
(np = 100, nt = 50, r_par in [-200, 200])."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import bench  # noqa: E402
from picca_b200 import synth, xcf  # noqa: E402
from picca_b200.engine import get_engine  # noqa: E402
from tests import helpers  # noqa: E402

nq = int(sys.argv[1]) if len(sys.argv) > 1 else 170000
workload = sys.argv[2] if len(sys.argv) > 2 else "c2_100k"
t0 = time.time()
kw = dict(bench.WORKLOADS[workload])
n = kw.pop("n_forest")
data, num, z_min, z_max, cosmo = synth.make_forests(n, **kw)
objs, z_min2 = synth.make_quasars(nq, seed=20260103, nside=32, ra_deg=kw["ra_deg"],
                                  dec_deg=kw["dec_deg"], cosmo=cosmo)
print("generated %d forests and %d quasars in %.1fs" % (num, nq, time.time() - t0), flush=True)
ang_max = synth.compute_ang_max(cosmo, 200., z_min, z_min2)
helpers.configure(xcf, data, num, ang_max, objs=objs, num_bins_r_par=100, num_bins_r_trans=50,
                  r_par_max=200., r_par_min=-200., r_trans_max=200., nside=32, alpha_obj=1.44)
eng = get_engine()
torch = eng.torch
eng.lib.pb2_set_timing(1)
hps = sorted(data)
for rep in range(2):
    torch.cuda.synchronize()
    t0 = time.time()
    xcf.fill_neighs(hps)
    torch.cuda.synchronize()
    t1 = time.time()
    out = xcf.compute_xi_batch(hps, normalise=False)
    torch.cuda.synchronize()
    t2 = time.time()
    pairs = int(out[:, 5, :].view(np.int64).sum())
    kms = eng.lib.pb2_last_kernel_ms()
    print("c3 xcf rep %d: %d forests x %d quasars, %d binned pairs, neighbours %.2fs, xi %.2fs wall, "
          "kernel %.1f ms -> %.3e pairs/s" % (rep, num, nq, pairs, t1 - t0, t2 - t1, kms,
                                              pairs / (kms * 1e-3)), flush=True)
