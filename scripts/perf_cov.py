"""Covariance step at the config-2 shape (1392 HEALPix sub-samples x 2500 bins; and 5000 bins):
kernel times (CUDA events on the launching stream, pb2_last_kernel_ms), FP64 roofline fraction
(unique DFMAs nb(nb+1)/2 x n_s over the measured DFMA issue peak) and the NumPy restatement
(oracle, host BLAS threads) beside it."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from oracle import export as oexp  # noqa: E402  (checker / CPU baseline only)
from picca_b200 import export  # noqa: E402
from picca_b200.engine import get_engine  # noqa: E402
from tests.golden import cases_export  # noqa: E402

export.userprint = lambda *a, **k: None
eng = get_engine()
torch = eng.torch
peak, _ = eng.fp64_peak(8192)
for n_s, np_, nt in ((1392, 50, 50), (1392, 100, 50)):
    cfg = dict(n_s=n_s, np_=np_, nt=nt, delta_r_par=4., delta_r_trans=4., seed=11)
    xi, we, rp, rt = cases_export.inputs(cfg)
    nb = np_ * nt
    d_xi, d_we = export._dev(eng, xi), export._dev(eng, we)
    eng.lib.pb2_set_timing(1)
    ms = []
    for rep in range(6):
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=eng.device).fill_(rep)  # > L2
        cov, _, _ = export.compute_cov_device(eng, d_xi, d_we)
        torch.cuda.synchronize()
        ms.append(eng.lib.pb2_last_kernel_ms())
    kms = float(np.median(ms[2:]))
    ops = nb * (nb + 1) / 2. * n_s
    t0 = time.perf_counter()
    host = export.compute_cov(xi, we)
    e2e = time.perf_counter() - t0
    t0 = time.perf_counter()
    want = oexp.compute_cov(xi, we)
    cpu = time.perf_counter() - t0
    sd = np.sqrt(np.diagonal(want))
    err = np.max(np.abs(host - want) / np.maximum(sd[:, None] * sd[None, :], 1e-300))
    print("cov n_s=%d nb=%d: kernels %.3f ms (prepare + syrk), %.2f Tops/s = %.3f of the DFMA "
          "peak %.2f; host-to-host %.1f ms; numpy (%d threads) %.1f ms; max scaled err %.2e"
          % (n_s, nb, kms, ops / kms / 1e9, ops / (kms * 1e-3) / peak, peak / 1e12, e2e * 1e3,
             os.cpu_count(), cpu * 1e3, err), flush=True)
    ms = []
    d_cov = export._dev(eng, want)
    for rep in range(4):
        t0 = time.perf_counter()
        smooth = export.smooth_cov(None, None, rp, rt, covariance=want)
        ms.append((time.perf_counter() - t0) * 1e3)
        kms = eng.lib.pb2_last_kernel_ms()
    t0 = time.perf_counter()
    if nb <= 2500:
        want_s = oexp.smooth_cov(None, None, rp, rt, covariance=want)
        cpu = time.perf_counter() - t0
        err = np.max(np.abs(smooth - want_s) / np.maximum(sd[:, None] * sd[None, :], 1e-300))
    else:
        cpu, err = float("nan"), float("nan")
    print("smooth nb=%d: kernels %.3f ms, host-to-host %.1f ms; numpy restatement %.1f ms "
          "(the reference's Python double loop visits %d bin pairs twice); max scaled err %.2e"
          % (nb, kms, float(np.median(ms[1:])), cpu * 1e3, nb * (nb - 1) // 2, err), flush=True)
    eng.lib.pb2_set_timing(0)

# ---- bootstrap covariance (picca_export.py --num-boot-cov, utils.py:131-150), default 10000
cfg = dict(n_s=1392, np_=50, nt=50, delta_r_par=4., delta_r_trans=4., seed=11)
xi, we, _, _ = cases_export.inputs(cfg)
eng.lib.pb2_set_timing(1)
for rep in range(2):
    t0 = time.perf_counter()
    boot = export.compute_cov_boot(xi, we, nboots=10000)
    dt = time.perf_counter() - t0
    kms = eng.lib.pb2_last_kernel_ms()
eng.lib.pb2_set_timing(0)
ops = 2. * 10000 * 1392 * 2500 + 2500 * 2501 / 2. * 10000
t0 = time.perf_counter()
oexp.compute_cov_boot(xi, we, nboots=20)
cpu = (time.perf_counter() - t0) / 20.
print("boot 10000 x (1392 x 2500): kernels %.2f ms = %.2f Tops/s = %.3f of the DFMA peak; whole "
      "call %.2f s (host draw of the 10000 index sets + PCIe); the reference's loop costs %.1f ms "
      "per realisation on this host (20 timed) -> %.0f s for 10000"
      % (kms, ops / kms / 1e9, ops / (kms * 1e-3) / peak, dt, cpu * 1e3, cpu * 1e4), flush=True)
