"""Helpers that drive the UNMODIFIED reference (imported from /root/reference through the shims)
to load its bundled fixtures and run its own Numba path.  Test infrastructure, usable only where
/root/reference exists (this container; never on the GPU box)."""
import importlib
import warnings

from . import shims

DATA = shims.REFERENCE_DATA


def reference_modules():
    """(cf, xcf, io, constants, utils) of the live reference, freshly reloaded so that Numba
    re-reads the module globals."""
    assert shims.install(), "reference not available"
    warnings.filterwarnings("ignore")
    import picca.cf, picca.xcf, picca.io, picca.constants, picca.utils  # noqa: E401
    cf = importlib.reload(picca.cf)
    xcf = importlib.reload(picca.xcf)
    return cf, xcf, picca.io, picca.constants, picca.utils


def quiet(mod):
    mod.userprint = lambda *a, **k: None


class DummyLock:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


class DummyCounter:
    value = 0


def load_deltas(nside=16, in_dir=None, lambda_abs_name="LYA", z_evol=2.9, z_ref=2.25,
                fid_Om=0.315, nspec=None, no_project=False):
    """data, num_data, z_min, z_max, cosmo -- exactly what picca_cf.py builds (:371-405)."""
    _, _, io, constants, _ = reference_modules()
    quiet(io)
    in_dir = in_dir or (DATA + "/test_delta/Delta_LYA/")
    cosmo = constants.Cosmo(Om=fid_Om, Or=0., Ok=0., wl=-1., blinding="none")
    lambda_abs = constants.ABSORBER_IGM[lambda_abs_name]
    data, num_data, z_min, z_max = io.read_deltas(
        in_dir, nside, lambda_abs, z_evol, z_ref, cosmo, max_num_spec=nspec,
        no_project=no_project, nproc=1,
        delta_attributes=DATA + "/test_delta/delta_attributes.fits.gz")
    return data, num_data, z_min, z_max, cosmo


def load_objects(cosmo, nside=16, z_min_obj=0., z_max_obj=10., z_evol_obj=1., z_ref=2.25):
    """objs, z_min2 -- what picca_xcf.py builds (:403-412)."""
    _, _, io, _, _ = reference_modules()
    quiet(io)
    objs, z_min2 = io.read_objects(DATA + "/test_delta/cat.fits", nside, z_min_obj, z_max_obj,
                                   z_evol_obj, z_ref, cosmo, mode="sdss")
    return objs, z_min2


def to_reference_deltas(data):
    """Rebuild a dict of picca_b200.forest.Delta stand-ins as the reference's own
    ``picca.data.Delta`` objects (so that its get_angle_between etc. are the code under test)."""
    assert shims.install()
    from picca.data import Delta
    out = {}
    for hp, forests in data.items():
        out[hp] = []
        for d in forests:
            r = Delta(d.thingid, d.ra, d.dec, d.z_qso, d.plate, d.mjd, d.fiberid,
                      d.log_lambda.copy(), d.weights.copy(), None, d.delta.copy(), d.order,
                      None, None, None, None, None)
            r.z, r.r_comov, r.dist_m = d.z.copy(), d.r_comov.copy(), d.dist_m.copy()
            out[hp].append(r)
    return out


def to_reference_qsos(objs):
    assert shims.install()
    from picca.data import QSO
    out = {}
    for hp, qsos in objs.items():
        out[hp] = []
        for q in qsos:
            r = QSO(q.thingid, q.ra, q.dec, q.z_qso, q.plate, q.mjd, q.fiberid)
            r.weights, r.r_comov, r.dist_m = q.weights, q.r_comov, q.dist_m
            r.log_lambda = q.log_lambda
            out[hp].append(r)
    return out
