"""Overlay package: makes ``picca.cf`` and ``picca.xcf`` resolve to the B200 implementation while
every other ``picca.*`` submodule (io, data, constants, utils, prep_del, bin, ...) keeps coming
from the installed reference.  Put the parent directory of this package FIRST on ``sys.path``
(``picca_b200.overlay.activate()`` does it) and the reference's unmodified scripts --
picca_cf.py, picca_xcf.py, picca_dmat.py, picca_xdmat.py -- run on the GPU path.

The reference's ``picca/__init__.py`` is a regular package (py/picca/__init__.py:1-3), so a
namespace merge is not available; instead this package's ``__path__`` lists its own directory
first and the reference's ``picca`` directory second.
"""
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
__path__ = [_here]


def _reference_dir():
    env = os.environ.get("PICCA_REFERENCE_PATH")
    candidates = [env] if env else []
    candidates += list(sys.path)
    for entry in candidates:
        if not entry:
            continue
        cand = os.path.join(entry, "picca")
        if os.path.isfile(os.path.join(cand, "__init__.py")) and \
                os.path.realpath(cand) != os.path.realpath(_here):
            return cand
    return None


_ref = _reference_dir()
if _ref is not None:
    __path__.append(_ref)
    try:  # same version string as the reference package exposes
        with open(os.path.join(_ref, "_version.py")) as _f:
            exec(_f.read())
    except OSError:
        pass
