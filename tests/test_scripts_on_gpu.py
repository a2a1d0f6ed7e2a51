"""GPU: the reference's UNMODIFIED scripts (staged, unmodified, under the git-ignored
``baseline/_ref`` by scripts/stage_reference.py) run end to end on the B200 path through
``python -m picca_b200.run`` -- argument parsing, delta reading (B200 loader, isolated in a child
process), fork pools, FITS writing are the reference's own code; only ``picca.cf`` / ``picca.xcf``
/ ``picca.io.read_deltas`` resolve to picca_b200.  Exact command lines of the reference's
tests/test_3_cor.py (:206-257, :318-349, :611-698) on its bundled ``Delta_LYA`` fixtures; outputs
compared with its golden FITS by its own ``compare_fits`` rule (array_equal, else
allclose(rtol=1e-5, atol=1e-8)) PLUS bit-equal ``NB`` / ``NPALL`` / ``NPUSED``.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

from tests.refharness import shims

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not shims.reference_available(),
                                 reason="needs the staged reference (baseline/_ref)")]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DATA = shims.REFERENCE_DATA or ""
DELTAS = (" --in-attributes " + DATA + "/test_delta/delta_attributes.fits.gz --in-dir " + DATA +
          "/test_delta/Delta_LYA/")
COMMON = " --rp-max +60.0 --rt-max +60.0 --nt 15" + DELTAS
DRQ = " --drq " + DATA + "/test_delta/cat.fits"


def run_script(impl, script, flags, out, env=None):
    cmd = [sys.executable, "-m", "tests.refharness.run_script", "--impl", impl, script] + \
        (flags + " --out " + out).split()
    full_env = dict(os.environ, **(env or {}))
    res = subprocess.run(cmd, cwd=ROOT, env=full_env, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, "%s failed:\n%s\n%s" % (script, res.stdout[-3000:],
                                                         res.stderr[-3000:])
    return res.stdout


def read_fits(path):
    from tests.refharness import minifits
    return minifits.FITS(path)


def compare_fits(path_got, path_want, exact=()):
    """reference tests/test_helpers.py:45-112 + bit-equality of the columns in ``exact``"""
    got, want = read_fits(path_got), read_fits(path_want)
    assert len(got) == len(want)
    for h in range(1, len(want)):
        tg, tw = got[h].read(), want[h].read()
        assert sorted(tg.dtype.names) == sorted(tw.dtype.names)
        for name in tw.dtype.names:
            if name in exact:
                assert np.array_equal(tg[name], tw[name]), (h, name)
            elif not np.array_equal(tg[name], tw[name]):
                assert np.allclose(tg[name], tw[name], rtol=1e-5, atol=1e-8), (h, name)


@pytest.mark.parametrize("script,flags,golden,nproc", [
    ("picca_cf.py", COMMON + " --rp-min +0.0 --np 15 --remove-same-half-plate-close-pairs", "cf", 2),
    ("picca_cf.py", COMMON + " --rp-min +0.0 --np 15 --remove-same-half-plate-close-pairs", "cf", 1),
    ("picca_xcf.py", COMMON + " --rp-min -60.0 --np 30 --z-evol-obj 1." + DRQ, "xcf", 2),
    ("picca_cf_angl.py", DELTAS, "cf_angl", 2),
    ("picca_xcf_angl.py", DELTAS + " --z-evol-obj 1." + DRQ, "xcf_angl", 2),
])
def test_correlation_script_matches_reference_golden(tmp_path, script, flags, golden, nproc):
    out = str(tmp_path / (golden + ".fits.gz"))
    run_script("b200", script, flags + " --nproc %d" % nproc, out)
    compare_fits(out, DATA + "/test_cor/" + golden + ".fits.gz", exact=("NB", "HEALPID"))


@pytest.mark.parametrize("script,flags,golden", [
    ("picca_dmat.py", COMMON + " --rp-min +0.0 --np 15 --rej 0.99 --nproc 1"
     " --remove-same-half-plate-close-pairs --no-redshift-evolution", "dmat"),
    ("picca_xdmat.py", COMMON + " --rp-min -60.0 --np 30 --rej 0.99 --nproc 1 --z-evol-obj 1."
     " --no-redshift-evolution" + DRQ, "xdmat"),
])
def test_distortion_script_matches_reference_golden(tmp_path, script, flags, golden):
    out = str(tmp_path / (golden + ".fits.gz"))
    run_script("b200", script, flags, out)
    want = DATA + "/test_cor/" + golden + ".fits.gz"
    compare_fits(out, want)
    hg, hw = read_fits(out)[1].read_header(), read_fits(want)[1].read_header()
    assert (hg["NPALL"], hg["NPUSED"]) == (hw["NPALL"], hw["NPUSED"])


@pytest.mark.parametrize("script,flags,golden", [
    # test_3_cor.py:413-445, :759-793 (default --max-diagram: T1-T3 / T1-T4)
    ("picca_wick.py", COMMON + " --rp-min +0.0 --np 15 --rej 0.99 --nproc 1 --cf1d " + DATA +
     "/test_cor/cf1d.fits.gz", "wick"),
    ("picca_xwick.py", COMMON + " --rp-min -60.0 --np 30 --rej 0.99 --nproc 1 --z-evol-obj 1."
     " --cf1d " + DATA + "/test_cor/cf1d.fits.gz" + DRQ, "xwick"),
])
def test_wick_script_matches_reference_golden(tmp_path, script, flags, golden):
    out = str(tmp_path / (golden + ".fits.gz"))
    run_script("b200", script, flags, out)
    compare_fits(out, DATA + "/test_cor/" + golden + ".fits.gz", exact=("NB",))


def test_dmat_script_fork_pool_matches_oracle_with_same_chunking(tmp_path):
    """--nproc 2: two forked workers, round-robin HEALPix chunks, one seed per chunk
    (picca_dmat.py:36, :471-501).  The result depends on --nproc, so the comparison is
    with the same unmodified script driving the oracle double (CPU) at the same --nproc."""
    flags = COMMON + " --rp-min +0.0 --np 15 --rej 0.95 --nproc 2"
    got, want = str(tmp_path / "gpu.fits.gz"), str(tmp_path / "cpu.fits.gz")
    run_script("b200", "picca_dmat.py", flags, got)
    run_script("oracle", "picca_dmat.py", flags, want)
    fg, fw = read_fits(got), read_fits(want)
    hg, hw = fg[1].read_header(), fw[1].read_header()
    assert (hg["NPALL"], hg["NPUSED"]) == (hw["NPALL"], hw["NPUSED"]) and hw["NPUSED"] > 100
    tg, tw = fg[1].read(), fw[1].read()
    scale = np.abs(tw["DM"]).max()
    assert np.abs(tg["DM"] - tw["DM"]).max() <= 1e-9 * scale   # north_star: 1e-9 relative
    np.testing.assert_allclose(tg["WDM"], tw["WDM"], rtol=1e-9)
