"""Host side of the Wick-expansion kernels (pb2_wick.cu): the per-pixel inputs the reference
evaluates per forest with NumPy / scipy -- get_variance_1d(log_lambda), the redshift-evolution
factors, the xi_1d interpolator (reference py/picca/cf.py:1412-1421, :1556-1557;
picca_wick.py:393-417) -- evaluated once per catalogue and placed in HBM."""
import numpy as np


def nearest_table(fn):
    """(bounds, values) of a ``scipy.interpolate.interp1d(kind="nearest",
    fill_value="extrapolate")`` -- what picca_wick.py / picca_xwick.py build for xi_1d
    (picca_wick.py:412-417): the interpolator returns ``y[searchsorted(x_bds, x, "left").clip(0,
    n-1)]`` with ``x_bds = x[1:]/2 + x[:-1]/2``; the device does the same search on the same
    bounds.  Other callables cannot be evaluated on the device: NotImplementedError (loud)."""
    kind = getattr(fn, "_kind", None)
    if kind != "nearest" or not hasattr(fn, "x") or not hasattr(fn, "y") or \
            getattr(fn, "_side", "left") != "left":
        raise NotImplementedError(
            "picca_b200: xi_1d must be a scipy interp1d(kind='nearest') as picca_wick.py builds "
            "it (picca_wick.py:412-417); got %r" % (fn,))
    x = np.asarray(fn.x, dtype=np.float64)
    y = np.ascontiguousarray(np.asarray(fn.y, dtype=np.float64).reshape(-1))
    if hasattr(fn, "x_bds"):
        bounds = np.ascontiguousarray(np.asarray(fn.x_bds, dtype=np.float64))
    else:
        half = x / 2.0
        bounds = np.ascontiguousarray(half[1:] + half[:-1])
    if y.size != x.size or bounds.size != x.size - 1:
        raise NotImplementedError("picca_b200: xi_1d table with a vector-valued y")
    return bounds, y


def catalogue_fname(host):
    """The ``fname`` label of a delta catalogue (picca_wick.py:371, :475: "D1" / "D2", one per
    catalogue)."""
    names = {getattr(o, "fname", None) for o in host.objs}
    if len(names) != 1 or None in names:
        raise RuntimeError("picca_b200: every delta of a catalogue must carry the same `fname` "
                           "(picca_wick.py:371); found %r" % (sorted(map(str, names)),))
    return names.pop()


def pixel_inputs(eng, host, get_variance_1d, xi_1d, z_ref, alpha):
    """Device tensors (variance_1d per pixel, evolution factor per pixel, xi_1d bounds, xi_1d
    values) of one delta catalogue."""
    torch = eng.torch
    fname = catalogue_fname(host)
    ll, z = host.arrays["log_lambda"], host.arrays["z"]
    var = np.ascontiguousarray(np.asarray(get_variance_1d[fname](ll), dtype=np.float64))
    ze = np.ascontiguousarray(((1 + z) / (1 + z_ref))**(alpha - 1))   # cf.py:1556
    bounds, values = nearest_table(xi_1d[fname])
    up = lambda a: torch.from_numpy(a).to(eng.device)
    return up(var), up(ze), up(bounds if bounds.size else np.zeros(1)), up(values), values.size
