#!/usr/bin/env python
"""Stage the UNMODIFIED reference under the git-ignored ``baseline/_ref`` so that it travels to
the GPU box with the gpurun snapshot (the box has no /root/reference):

    python scripts/stage_reference.py [--force]

1. ``pip install --no-index --no-build-isolation --no-deps --target baseline/_ref`` of a /tmp copy
   of /root/reference (the build writes egg-info into the source tree; /root/reference is
   read-only).  Installs the ``picca`` package incl. ``picca/bin/picca_cf.py`` etc.
2. The reference's wheel carries no test data: the bundled delta fixtures and the golden
   correlation FITS the parity tests compare with (``py/picca/tests/data/test_delta``,
   ``test_cor``) are copied next to the installed package (``baseline/_ref/picca/tests/data``).

Nothing under ``baseline/_ref`` is tracked by git or imported by ``picca_b200``: it is the
reference arm of ``bench.py`` and the input of the script-level GPU tests
(``tests/test_scripts_on_gpu.py``).
"""
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = "/root/reference"
TARGET = os.path.join(ROOT, "baseline", "_ref")
DATA_ITEMS = [
    "test_delta/Delta_LYA", "test_delta/Delta_LYA_image", "test_delta/cat.fits",
    "test_delta/delta_attributes.fits.gz", "test_delta/random.fits", "test_cor",
]


def staged():
    return os.path.isfile(os.path.join(TARGET, "picca", "cf.py")) and \
        os.path.isdir(os.path.join(TARGET, "picca", "tests", "data", "test_cor"))


def stage(force=False):
    """Returns the target directory, or None when /root/reference is absent."""
    if staged() and not force:
        return TARGET
    if not os.path.isdir(os.path.join(REFERENCE, "py", "picca")):
        return None
    shutil.rmtree(TARGET, ignore_errors=True)
    os.makedirs(TARGET, exist_ok=True)
    with tempfile.TemporaryDirectory(prefix="picca_src_") as tmp:
        src = os.path.join(tmp, "picca")
        shutil.copytree(REFERENCE, src, ignore=shutil.ignore_patterns("tests", ".git"))
        # the package list is discovered from the tree: keep picca.tests importable but empty
        os.makedirs(os.path.join(src, "py", "picca", "tests"), exist_ok=True)
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation",
               "--no-deps", "--find-links", "/opt/wheelhouse", "--target", TARGET, src]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError("pip install of the reference into baseline/_ref failed")
    data_src = os.path.join(REFERENCE, "py", "picca", "tests", "data")
    data_dst = os.path.join(TARGET, "picca", "tests", "data")
    for item in DATA_ITEMS:
        s, d = os.path.join(data_src, item), os.path.join(data_dst, item)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        if os.path.isdir(s):
            shutil.copytree(s, d, ignore=shutil.ignore_patterns("input_from_delta_extraction*"))
        else:
            shutil.copy2(s, d)
    return TARGET


if __name__ == "__main__":
    out = stage(force="--force" in sys.argv)
    print(out if out else "reference not available here; baseline/_ref not staged")
