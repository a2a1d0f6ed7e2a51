"""Metal distortion matrix at the BASELINE config-4 shape on a reduced footprint: --rej 0.99,
np = nt = 50, LYA x SiIII(1207) (both passes) on the synthetic 20k-forest sample; kernel time,
used forest pairs/s and contributing pixel pairs/s, with the oracle (the reference's NumPy
algorithm, 1 core) on a few hundred kept forest pairs beside it."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import bench  # noqa: E402
from picca_b200 import cf  # noqa: E402
from picca_b200.engine import get_engine  # noqa: E402
from tests import helpers  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "c2_20k"
data, num, ang_max = bench.make_workload(workload)[:3]
from picca_b200 import synth  # noqa: E402
cosmo = synth.FlatLCDM()
cfg = dict(num_bins_r_par=50, num_bins_r_trans=50, r_par_max=200., r_trans_max=200.,
           num_model_bins_r_par=50, num_model_bins_r_trans=50, nside=32, reject=0.99)
helpers.configure(cf, data, num, ang_max, **cfg)
cf.alpha_abs, cf.cosmo = {"LYA": 2.9, "SiIII(1207)": 1.}, cosmo
eng = get_engine()
eng.lib.pb2_set_timing(1)
hps = sorted(data)
for rep in range(2):
    eng.torch.cuda.synchronize()
    t0 = time.time()
    cf.fill_neighs(hps)
    np.random.seed(hps[0])
    res = cf.compute_metal_dmat(hps, "LYA", "SiIII(1207)")
    eng.torch.cuda.synchronize()
    dt = time.time() - t0
    kms = eng.lib.pb2_last_kernel_ms()
    print("metal dmat rep %d (%s): NPALL %d NPUSED %d, %.2fs wall, last pass kernel %.1f ms, "
          "sum(dmat)=%.6e sum(weights_dmat)=%.6e" % (rep, workload, res[6], res[7], dt, kms,
                                                     res[1].sum(), res[0].sum()), flush=True)
eng.lib.pb2_set_timing(0)

from oracle import cf as ocf  # noqa: E402  (CPU baseline only)
helpers.configure(ocf, data, num, ang_max, **cfg)
ocf.alpha_abs, ocf.cosmo = cf.alpha_abs, cosmo
sample = hps[len(hps) // 2:len(hps) // 2 + 2]
ocf.fill_neighs(sample)
np.random.seed(sample[0])
t0 = time.time()
o = ocf.compute_metal_dmat(sample, "LYA", "SiIII(1207)")
dt = time.time() - t0
print("oracle (NumPy, 1 core): %d used forest pairs in %.2fs -> %.1f used forest pairs/s"
      % (o[7], dt, o[7] / dt))
