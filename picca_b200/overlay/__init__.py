"""Drop-in activation of the B200 path under the reference's unmodified scripts."""
import os
import sys

OVERLAY_DIR = os.path.dirname(os.path.abspath(__file__))


def activate(reference_py=None):
    """Put the overlay ``picca`` package ahead of the reference on ``sys.path``.
    ``reference_py``: directory that contains the reference's ``picca`` package (optional when it
    is already importable)."""
    if reference_py:
        os.environ["PICCA_REFERENCE_PATH"] = reference_py
        if reference_py not in sys.path:
            sys.path.append(reference_py)
    for name in [m for m in sys.modules if m == "picca" or m.startswith("picca.")]:
        del sys.modules[name]
    if OVERLAY_DIR in sys.path:
        sys.path.remove(OVERLAY_DIR)
    sys.path.insert(0, OVERLAY_DIR)
