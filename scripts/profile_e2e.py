"""Host-side profile (cProfile) of one end-to-end auto-correlation call on host data:
catalog pack + H2D + neighbours + kernel + D2H, as bench.py's `e2e` arm at N = 1 does it.
usage: python scripts/profile_e2e.py [workload]"""
import cProfile
import pstats
import sys
import time

sys.path.insert(0, ".")
import bench  # noqa: E402
from picca_b200 import catalog, cf  # noqa: E402
from picca_b200.engine import get_engine  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "c2_100k"
data, num, ang_max, cosmo, z_min = bench.make_workload(workload)
bench.configure(cf, data, num, ang_max)
cf.userprint = lambda *a, **k: None
eng = get_engine()
torch = eng.torch
hps = sorted(data)


def step():
    catalog.invalidate(data)
    eng.drop_catalogs()
    catalog.cached_pack(data)
    cf.fill_neighs(hps)
    out = cf.compute_xi_batch(hps)
    torch.cuda.synchronize()
    return out


step()
eng.lib.pb2_set_timing(1)
t0 = time.perf_counter()
step()
wall = time.perf_counter() - t0
print("e2e step %.3f s wall, pair kernel %.3f s -> host + transfer overhead %.3f s"
      % (wall, eng.lib.pb2_last_kernel_ms() * 1e-3, wall - eng.lib.pb2_last_kernel_ms() * 1e-3))
eng.lib.pb2_set_timing(0)
pr = cProfile.Profile()
pr.enable()
step()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(45)
