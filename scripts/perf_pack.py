"""Device-side packing of a catalogue at config-2 size: host pack(), H2D of the SoA, pb2_pack_diag
and pb2_build_prefix (CUDA events on the current stream)."""
import ctypes
import sys
import time

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from picca_b200 import _lib, catalog  # noqa: E402
from picca_b200.engine import get_engine  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "c2_100k"
data, num, ang_max = bench.make_workload(workload)[:3]
t0 = time.perf_counter()
host = catalog.pack(data)
print("host pack(): %.2f s for %d forests, %d pixels, %.2f GB of host arrays"
      % (time.perf_counter() - t0, host.n_los, host.n_pix, host.nbytes() / 1e9), flush=True)
eng = get_engine()
pinned = {k: torch.from_numpy(v).pin_memory() for k, v in host.arrays.items()}
for rep in range(3):
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    ev[0].record()
    tensors = {k: v.to(eng.device, non_blocking=True) for k, v in pinned.items()}
    ev[1].record()
    dev = catalog.DeviceCatalog.from_tensors(host, eng.device, tensors)
    ev[2].record()
    torch.cuda.synchronize()
    rec_bytes = (dev.tensors["dg_rec"].numel() + dev.tensors["il_rec"].numel() +
                 dev.tensors["px_rec"].numel()) * 8
    ms = ev[1].elapsed_time(ev[2])
    print("rep %d: H2D of the SoA %.1f ms (%.2f GB); pb2_pack_diag + pb2_build_prefix %.2f ms for "
          "%.2f GB of derived records (%.0f GB/s written); %.2f GB in HBM"
          % (rep, ev[0].elapsed_time(ev[1]), host.nbytes() / 1e9, ms, rec_bytes / 1e9,
             rec_bytes / ms / 1e6, dev.device_bytes() / 1e9), flush=True)
    del dev, tensors
