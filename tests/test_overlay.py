"""CPU, this container only: the overlay makes the reference's UNMODIFIED scripts bind the B200
modules as picca.cf / picca.xcf while everything else stays the reference's, and the product
fails loudly (no CPU fallback) when no CUDA device is present."""
import importlib
import sys

import pytest

from tests.refharness import shims

pytestmark = [pytest.mark.reference,
              pytest.mark.skipif(not shims.reference_available(), reason="needs /root/reference")]


@pytest.fixture()
def overlay_on():
    shims.install()
    from picca_b200 import overlay
    saved = {k: v for k, v in sys.modules.items() if k == "picca" or k.startswith("picca.")}
    saved_path = list(sys.path)
    overlay.activate(shims.REFERENCE_PY)
    yield
    for name in [m for m in sys.modules if m == "picca" or m.startswith("picca.")]:
        del sys.modules[name]
    sys.modules.update(saved)
    sys.path[:] = saved_path


def test_scripts_bind_b200_modules(overlay_on):
    import picca_b200.cf
    import picca_b200.xcf
    cf_script = importlib.import_module("picca.bin.picca_cf")
    dmat_script = importlib.import_module("picca.bin.picca_dmat")
    xcf_script = importlib.import_module("picca.bin.picca_xcf")
    xdmat_script = importlib.import_module("picca.bin.picca_xdmat")
    assert cf_script.cf is picca_b200.cf and dmat_script.cf is picca_b200.cf
    assert xcf_script.xcf is picca_b200.xcf and xdmat_script.xcf is picca_b200.xcf
    import picca.io
    import picca.constants
    assert picca.io.__file__.startswith(shims.REFERENCE_PY)
    assert picca.constants.__file__.startswith(shims.REFERENCE_PY)
    # reload (used by the reference's tests, test_3_cor.py:233) keeps working
    again = importlib.reload(sys.modules["picca.cf"])
    assert again.compute_xi is not None


def test_exports_every_global_and_function_of_the_reference(overlay_on):
    """every module global of reference cf.py:28-79 / xcf.py:27-68 and the hot-path functions"""
    import ast
    import picca_b200.cf
    import picca_b200.xcf
    for name, mod in (("cf", picca_b200.cf), ("xcf", picca_b200.xcf)):
        tree = ast.parse(open(shims.REFERENCE_PY + "/picca/%s.py" % name).read())
        for node in tree.body:
            if isinstance(node, ast.Assign):
                for tgt in node.targets:
                    assert hasattr(mod, tgt.id), (name, tgt.id)
        for fn in ("fill_neighs", "compute_xi", "compute_dmat", "compute_xi_forest_pairs_fast",
                   "compute_dmat_forest_pairs_fast"):
            assert callable(getattr(mod, fn)), (name, fn)


def test_product_fails_loudly_without_cuda(overlay_on):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import picca_b200.cf as cf
    from tests import helpers
    data, num, z_min, _, cosmo = helpers.small_sample(n=20, seed=2, max_pix=30)
    helpers.configure(cf, data, num, 0.01)
    hps = sorted(data)[:1]
    # in a CUDA-free main process fill_neighs is deferred (picca_xwick.py fills before it forks
    # its pool); the first use of the neighbours, or compute_*, needs the device: loud failure
    cf.fill_neighs(hps)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        len(data[hps[0]][0].neighbours)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        cf.compute_xi(hps)


def test_export_script_binds_b200_covariance(overlay_on):
    """picca_export.py does `from picca.utils import smooth_cov, compute_cov` (picca_export.py:17):
    through the overlay these are picca_b200.export's, the rest of picca.utils is the reference's
    own code."""
    import picca_b200.export
    script = importlib.import_module("picca.bin.picca_export")
    assert script.compute_cov is picca_b200.export.compute_cov
    assert script.smooth_cov is picca_b200.export.smooth_cov
    import picca.utils
    assert picca.utils.__file__.startswith(shims.REFERENCE_PY)
    assert callable(picca.utils.reference_compute_cov) and callable(picca.utils.compute_ang_max)
    # no CUDA device in this container: the product path must raise, never fall back
    import numpy as np
    with pytest.raises(RuntimeError):
        script.compute_cov(np.ones((3, 4)), np.ones((3, 4)))


def test_scripts_bind_b200_loader(overlay_on):
    """picca_cf.py calls io.read_deltas (picca_cf.py:387-405): through the overlay that is the
    B200 loader, the rest of picca.io is the reference's own code; without a CUDA device it
    raises instead of falling back."""
    cf_script = importlib.import_module("picca.bin.picca_cf")
    import picca.io
    assert cf_script.io is picca.io
    assert picca.io.__file__.startswith(shims.REFERENCE_PY)
    assert picca.io.read_deltas is not picca.io.reference_read_deltas
    assert callable(picca.io.read_objects)
    with pytest.raises(RuntimeError):
        picca.io.read_deltas(shims.REFERENCE_DATA + "/test_delta/Delta_LYA/",
                             16, 1215.67, 2.9, 2.25, None, delta_attributes=shims.REFERENCE_DATA +
                             "/test_delta/delta_attributes.fits.gz")


def test_co_script_binds_b200_module(overlay_on):
    import picca_b200.co
    script = importlib.import_module("picca.bin.picca_co")
    assert script.co is picca_b200.co
    import ast
    tree = ast.parse(open(shims.REFERENCE_PY + "/picca/co.py").read())
    for node in tree.body:
        if isinstance(node, ast.Assign):
            for tgt in node.targets:
                assert hasattr(picca_b200.co, tgt.id), tgt.id
    assert callable(picca_b200.co.fill_neighs) and callable(picca_b200.co.compute_xi)


def test_loader_rejection_is_loud_unless_fallback_is_requested(overlay_on, monkeypatch):
    """No silent CPU fallback on the delta loader: an input the B200 loader
    rejects raises; only PICCA_B200_IO_FALLBACK=1 hands it to the reference's read_deltas."""
    import picca.io
    import picca_b200.io as impl

    def reject(*args, **kwds):
        raise NotImplementedError("picca_b200.io: mixed LOGLAM / LAMBDA delta files")
    monkeypatch.setattr(impl, "read_deltas", reject)
    called = []
    monkeypatch.setattr(picca.io, "reference_read_deltas", lambda *a, **k: called.append(1) or 7)
    monkeypatch.delenv("PICCA_B200_IO_FALLBACK", raising=False)
    with pytest.raises(NotImplementedError):
        picca.io.read_deltas("x", 16, 1215.67, 2.9, 2.25, None)
    assert not called
    monkeypatch.setenv("PICCA_B200_IO_FALLBACK", "1")
    assert picca.io.read_deltas("x", 16, 1215.67, 2.9, 2.25, None) == 7 and called
    monkeypatch.setenv("PICCA_B200_IO", "0")  # explicit opt-out of the B200 loader
    assert picca.io.read_deltas("x", 16, 1215.67, 2.9, 2.25, None) == 7 and len(called) == 2
