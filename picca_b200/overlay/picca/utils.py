"""``picca.utils`` = the reference's module with the covariance step of ``picca_export.py``
(``compute_cov``, ``compute_cov_boot``, ``smooth_cov``; py/picca/utils.py:100-150, :153-249) replaced by the B200 path.
Everything else in the module is the reference's own code, executed from its file."""
import importlib.util
import os
import sys

import picca as _pkg

_ref_file = None
for _d in list(_pkg.__path__)[1:]:
    if os.path.isfile(os.path.join(_d, "utils.py")):
        _ref_file = os.path.join(_d, "utils.py")
        break
if _ref_file is None:
    raise ImportError("picca_b200 overlay: the reference's picca/utils.py was not found")
_spec = importlib.util.spec_from_file_location(__name__, _ref_file)
_mod = importlib.util.module_from_spec(_spec)
sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)

import picca_b200.export as _impl  # noqa: E402

_mod.reference_compute_cov = _mod.compute_cov
_mod.reference_smooth_cov = _mod.smooth_cov
_mod.compute_cov = _impl.compute_cov
_mod.reference_compute_cov_boot = _mod.compute_cov_boot
_mod.compute_cov_boot = _impl.compute_cov_boot
_mod.smooth_cov = _impl.smooth_cov
