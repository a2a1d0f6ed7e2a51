"""BASELINE config 4 at full size on one GPU: distortion matrix, --rej 0.99, on the synthetic
100k-forest sample (2500 x 2500 dmat).  Prints used forest pairs/s and the wall time of
cf.fill_neighs + cf.compute_dmat over all HEALPix pixels (one chunk, seed = first pixel, as
picca_dmat.py --nproc 1 does)."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import bench  # noqa: E402
from picca_b200 import cf  # noqa: E402
from picca_b200.engine import get_engine  # noqa: E402
from tests import helpers  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "c2_100k"
t0 = time.time()
data, num, ang_max = bench.make_workload(workload)[:3]
print("generated %d forests in %.1fs" % (num, time.time() - t0), flush=True)
helpers.configure(cf, data, num, ang_max, num_bins_r_par=50, num_bins_r_trans=50, r_par_max=200.,
                  r_trans_max=200., num_model_bins_r_par=50, num_model_bins_r_trans=50, nside=32,
                  reject=0.99)
eng = get_engine()
torch = eng.torch
eng.lib.pb2_set_timing(1)
hps = sorted(data)
for rep in range(2):
    torch.cuda.synchronize()
    t0 = time.time()
    cf.fill_neighs(hps)
    np.random.seed(hps[0])
    res = cf.compute_dmat(hps)
    torch.cuda.synchronize()
    dt = time.time() - t0
    kms = eng.lib.pb2_last_kernel_ms()
    print("c4 dmat rep %d: %d forests, NPALL %d NPUSED %d, %.2fs wall (neighbours + draw + kernels + "
          "D2H), kernel %.1f ms -> %.1f used forest pairs/s, sum(dmat)=%.6e, sum(weights_dmat)=%.6e"
          % (rep, num, res[6], res[7], dt, kms, res[7] / (kms * 1e-3), res[1].sum(), res[0].sum()),
          flush=True)
