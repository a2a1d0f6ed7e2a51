"""``picca.io`` = the reference's module with ``read_deltas`` (py/picca/io.py:383-512) replaced
by the B200 delta loader.  Everything else in the module (read_objects, read_drq, read_blinding,
...) is the reference's own code, executed from its file.  Inputs the B200 loader does not
implement (it raises NotImplementedError: e.g. non-fp64 columns, unsorted wavelengths under
``rebin_factor``) are handed to the reference's own ``read_deltas`` with a notice; ``PICCA_B200_IO=0`` keeps the reference loader for everything."""
import importlib.util
import os
import sys

import picca as _pkg

_ref_file = None
for _d in list(_pkg.__path__)[1:]:
    if os.path.isfile(os.path.join(_d, "io.py")):
        _ref_file = os.path.join(_d, "io.py")
        break
if _ref_file is None:
    raise ImportError("picca_b200 overlay: the reference's picca/io.py was not found")
_spec = importlib.util.spec_from_file_location(__name__, _ref_file)
_mod = importlib.util.module_from_spec(_spec)
sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)

_mod.reference_read_deltas = _mod.read_deltas


def _read_deltas(*args, **kwds):
    if os.environ.get("PICCA_B200_IO", "1") != "1":
        return _mod.reference_read_deltas(*args, **kwds)
    import picca_b200.io as _impl
    try:
        return _impl.read_deltas(*args, **kwds)
    except NotImplementedError as err:
        _mod.userprint("picca_b200: %s -- using the reference's read_deltas for this input" % err)
        return _mod.reference_read_deltas(*args, **kwds)


_read_deltas.__doc__ = _mod.reference_read_deltas.__doc__
_mod.read_deltas = _read_deltas
