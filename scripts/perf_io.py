"""Delta loader at scale: N forests x ~500 pixels in BinTable delta files (uncompressed, so the
numbers are the loader's and not zlib's), read with picca_b200.io.read_deltas; stage breakdown and
the oracle loader (NumPy + scipy per forest, the reference's algorithm) on a sample beside it."""
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, ".")
from picca_b200 import io, synth  # noqa: E402
from picca_b200.engine import get_engine  # noqa: E402
from tests.refharness import minifits  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
PER_FILE = 500
io.userprint = lambda *a, **k: None
rng = np.random.default_rng(5)
tmp = tempfile.mkdtemp()
in_dir = os.path.join(tmp, "Delta")
os.makedirs(in_dir)
t0 = time.time()
primary = minifits._cards_to_bytes([("SIMPLE", True), ("BITPIX", 8), ("NAXIS", 0),
                                    ("EXTEND", True)])
raw_bytes = 0
for k in range(0, N, PER_FILE):
    out = bytearray(primary)
    for f in range(k, min(N, k + PER_FILE)):
        zq = float(rng.uniform(2.1, 3.5))
        lam = np.arange(1040. * (1 + zq), 1200. * (1 + zq), 0.8)[:700]
        n = lam.size
        head = [{"name": "RA", "value": float(rng.uniform(0., 2.))},
                {"name": "DEC", "value": float(rng.uniform(0., 0.7))},
                {"name": "Z", "value": zq}, {"name": "LOS_ID", "value": f + 1},
                {"name": "ORDER", "value": 1}]
        out += minifits._table_bytes([np.log10(lam), rng.normal(0, .3, n), rng.uniform(.5, 2, n),
                                      np.ones(n)], ["LOGLAM", "DELTA", "WEIGHT", "CONT"], None,
                                     head, str(f + 1))
    with open(os.path.join(in_dir, "delta-%d.fits" % (k // PER_FILE)), "wb") as fh:
        fh.write(bytes(out))
    raw_bytes += len(out)
attr = os.path.join(tmp, "delta_attributes.fits")
a = minifits.FITS(attr, "rw", clobber=True)
a.write([np.arange(2.)], names=["X"], header=[{"name": "FITORDER", "value": 1}],
        extname="FIT_METADATA")
a.close()
print("wrote %d forests, %.2f GB in %.1fs" % (N, raw_bytes / 1e9, time.time() - t0), flush=True)

cosmo = synth.FlatLCDM()
eng = get_engine()
kw = dict(nside=32, lambda_abs=synth.LYA, alpha=2.9, z_ref=2.25, cosmo=cosmo,
          delta_attributes=attr)
for rep in range(3):
    os.environ["PICCA_B200_IO_TIMING"] = "1" if rep == 2 else "0"
    if rep == 2:
        io.userprint = lambda *a, **k: print(*a, **k) if a and str(a[0]).startswith("picca_b200.io") else None
    launches = eng.launch_count()
    t0 = time.perf_counter()
    data, num, z_min, z_max = io.read_deltas(in_dir, **kw)
    dt = time.perf_counter() - t0
    npix = sum(len(d.weights) for v in data.values() for d in v)
    print("read_deltas rep %d: %d forests, %d pixels in %.2fs -> %.0f forests/s, %.2f GB/s of "
          "file bytes, %d kernel launches" % (rep, num, npix, dt, num / dt, raw_bytes / dt / 1e9,
                                              eng.launch_count() - launches), flush=True)

# the reference's algorithm (oracle restatement: per-HDU reads, per-forest NumPy + scipy) on 2 files
from oracle import io as oio  # noqa: E402  (CPU baseline only)
sample = os.path.join(tmp, "Sample")
os.makedirs(sample)
for k in range(2):
    os.symlink(os.path.join(in_dir, "delta-%d.fits" % k), os.path.join(sample, "delta-%d.fits" % k))
t0 = time.perf_counter()
_, n_s, _, _ = oio.read_deltas(sample, 32, synth.LYA, 2.9, 2.25, cosmo.table(),
                               delta_attributes=attr)
dt = time.perf_counter() - t0
print("oracle loader (1 core, %d forests): %.2fs -> %.0f forests/s" % (n_s, dt, n_s / dt))
