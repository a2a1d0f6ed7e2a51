"""``picca.io`` = the reference's module with ``read_deltas`` (py/picca/io.py:383-512) replaced
by the B200 delta loader.  Everything else in the module (read_objects, read_drq, read_blinding,
...) is the reference's own code, executed from its file.

The scripts call ``read_deltas`` in their parent process and fork the worker pool afterwards
(picca_cf.py:387, :455), and CUDA does not survive a fork: under the overlay the device part of
the loader therefore runs in a short-lived child process (``isolate=True``).

Environment: ``PICCA_B200_IO=0`` keeps the reference loader for everything (explicit opt-out).
Inputs the B200 loader does not implement raise ``NotImplementedError`` -- loudly; only with
``PICCA_B200_IO_FALLBACK=1`` are they handed to the reference's own ``read_deltas``."""
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
    if os.environ.get("PICCA_B200_IO", "1") == "0":
        return _mod.reference_read_deltas(*args, **kwds)
    import picca_b200.io as _impl
    kwds.setdefault("isolate", os.environ.get("PICCA_B200_IO_ISOLATE", "1") == "1")
    try:
        return _impl.read_deltas(*args, **kwds)
    except NotImplementedError as err:
        if os.environ.get("PICCA_B200_IO_FALLBACK", "0") != "1":
            raise
        kwds.pop("isolate")
        _mod.userprint("picca_b200: %s -- PICCA_B200_IO_FALLBACK=1: using the reference's "
                       "read_deltas for this input" % err)
        return _mod.reference_read_deltas(*args, **kwds)


_read_deltas.__doc__ = _mod.reference_read_deltas.__doc__
_mod.read_deltas = _read_deltas
