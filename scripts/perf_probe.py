"""Quick GPU probe: time the xi kernels on a dense synthetic patch at production binning and
cross-check the product kernel against the brute-force validation kernel."""
import argparse
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from picca_b200 import cf, synth  # noqa: E402
from picca_b200.engine import get_engine  # noqa: E402
from tests import helpers  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=3000)
ap.add_argument("--side", type=float, default=12.)
ap.add_argument("--nside", type=int, default=32)
ap.add_argument("--brute", type=int, default=1)
ap.add_argument("--reps", type=int, default=3)
args = ap.parse_args()

t = time.time()
data, num, z_min, z_max, cosmo = synth.make_forests(
    args.n, seed=5, nside=args.nside, ra_deg=(10., 10. + args.side), dec_deg=(5., 5. + args.side))
print("generated %d forests in %d healpix, %.1fs" % (num, len(data), time.time() - t))
ang_max = synth.compute_ang_max(cosmo, 200., z_min)
helpers.configure(cf, data, num, ang_max, num_bins_r_par=50, num_bins_r_trans=50, r_par_max=200.,
                  r_trans_max=200., nside=args.nside)
eng = get_engine()
torch = eng.torch
peak, _ = eng.fp64_peak(8192)
print("fp64 peak %.3e op/s" % peak)
hps = sorted(data)
eng.lib.pb2_set_timing(1)
res = {}
for variant in ([0, 3, 1] if args.brute == 1 else ([0, 3] if args.brute == 2 else [0])):
    cf._XI_VARIANT = variant
    for rep in range(args.reps):
        torch.cuda.synchronize()
        t0 = time.time()
        cf.fill_neighs(hps)
        torch.cuda.synchronize()
        t1 = time.time()
        out = cf.compute_xi_batch(hps, normalise=False)
        torch.cuda.synchronize()
        t2 = time.time()
        kms = eng.lib.pb2_last_kernel_ms()
        npairs = int(out[:, 5, :].view(np.int64).sum())
        print("variant %d rep %d: neigh %.3fs xi %.3fs kernel %.2f ms binned pairs %d -> %.3e pairs/s"
              " (kernel), roofline frac %.3f" % (variant, rep, t1 - t0, t2 - t1, kms, npairs,
                                                 npairs / (kms * 1e-3),
                                                 30. * npairs / (kms * 1e-3) / peak))
    res[variant] = out
if args.brute == 1:
    a, b = res[0], res[1]
    print("counts equal:", np.array_equal(a[:, 5].view(np.int64), b[:, 5].view(np.int64)))
    for k in range(5):
        d = np.abs(a[:, k] - b[:, k])
        print("field", k, "max rel", (d / np.maximum(np.abs(b[:, k]), 1e-300))[np.abs(b[:, k]) > 1e-6].max())
