"""B200 implementation of the hot path of ``picca.xcf`` behind the reference's module API.

Same module globals (reference py/picca/xcf.py:27-68), same functions, same return tuples:

    fill_neighs(healpixs)                         xcf.py:71-123
    compute_xi(healpixs) -> 6-tuple               xcf.py:126-220
    compute_dmat(healpixs) -> 8-tuple             xcf.py:325-424
    compute_xi_forest_pairs_fast(...)             xcf.py:223-322  (in-place accumulate)
    compute_dmat_forest_pairs_fast(...)           xcf.py:427-674  (in-place accumulate)

so that picca_xcf.py / picca_xdmat.py run unchanged once this module is importable as
``picca.xcf``.  All arithmetic runs in the CUDA kernels of libpicca_b200.so; no Numba, no CPU
fallback.  Globals are read at call time.
"""
import sys

import numpy as np

from . import _corr, catalog as _catalog
from .engine import MODE_XCF, get_engine
from .forest import Delta as _Delta, QSO as _QSO
from .params import params_from_module


def userprint(*args, **kwds):
    """reference py/picca/utils.py:20-28"""
    print(*args, **kwds)
    sys.stdout.flush()


# ---- module globals: names and defaults of reference xcf.py:27-68
num_bins_r_par = None
num_bins_r_trans = None
num_model_bins_r_par = None
num_model_bins_r_trans = None
r_par_max = None
r_par_min = None
r_trans_max = None
z_min_pairs = None
z_max_pairs = None
ang_max = None
nside = None

zerr_cut_deg = None
zerr_cut_kms = None

counter = None
num_data = None

z_ref = None
alpha = None
alpha_obj = None
lambda_abs = None
alpha_abs = None

data = None
objs = None

reject = None
lock = None

cosmo = None
rmu_binning = False
ang_correlation = False

# variables for distortion matrix
redshift_evolution_in_distortion_matrix = True

# variables used in the wick covariance matrix computation (kept for attribute parity)
get_variance_1d = {}
xi_1d = {}
max_diagram = None
xi_wick = None

_THIS = sys.modules[__name__]
_STORE = _corr.NeighbourStore()
_XI_VARIANT = 0


def _catalogs():
    eng, host1, dev1 = _corr.engine_and_catalog(data, ang_correlation=ang_correlation)
    _, host2, dev2 = _corr.engine_and_catalog(objs, is_object=True,
                                              ang_correlation=ang_correlation)
    return eng, host1, dev1, host2, dev2


def fill_neighs(healpixs):
    """Neighbouring objects of every forest of ``healpixs`` (xcf.py:71-123), incl. the optional
    quasar-pair zerr cut (:102-115) and the r_par pre-filter (:117-121), on the device."""
    healpixs = list(healpixs)
    if _corr.defer_fill():   # main process, no CUDA yet: see _corr.defer_fill
        _STORE.defer(healpixs, data, _fill_neighs_now)
        return
    _fill_neighs_now(healpixs)


def _fill_neighs_now(healpixs):
    healpixs = list(healpixs)
    eng, host1, dev1, host2, dev2 = _catalogs()
    params = params_from_module(_THIS, cross=True)
    index, ranges = _corr.forest_index_of(host1, healpixs)
    pairs = eng.neighbours(dev1, dev2, params, MODE_XCF, index)
    if _corr.HOST_ANGLES:
        _corr.apply_host_angles(pairs, host1, host2)
    _STORE.put(healpixs, pairs, ranges, (host1, host2))
    _corr.set_lazy_neighbours(host1.objs, index, pairs, host2.objs)


def _pairs_for(healpixs):
    _, host1, _, host2, _ = _catalogs()
    pairs = _STORE.take(healpixs, (host1, host2))
    if pairs is None:  # stored in different batches or for a re-packed catalogue: rebuild
        _fill_neighs_now(healpixs)
        pairs = _STORE.take(healpixs, (host1, host2))
    return pairs


def compute_xi(healpixs):
    """Cross-correlation of the forests of ``healpixs`` with their neighbouring objects
    (xcf.py:126-220).  Returns (weights, xi, r_par, r_trans, z, num_pairs), normalised per call."""
    healpixs = list(healpixs)
    eng, host1, dev1, host2, dev2 = _catalogs()
    params = params_from_module(_THIS, cross=True)
    pairs = _pairs_for(healpixs)
    out_row = eng.torch.zeros(pairs.n_f1, dtype=eng.torch.int32, device=eng.device)
    out = eng.xi(dev1, dev2, params, pairs, out_row, 1, cross_obj=True, variant=_XI_VARIANT,
                 normalise=True)
    host = out.cpu().numpy()[0]
    _corr.bump_progress(_THIS, pairs.n_f1, userprint)
    _corr.clear_neighbours(host1.objs, pairs.f1_index.cpu().numpy())  # xcf.py:213
    _STORE.drop(healpixs)
    weights, xi, r_par, r_trans, z = (np.ascontiguousarray(host[k]) for k in range(5))
    num_pairs = np.ascontiguousarray(host[5]).view(np.int64)
    return weights, xi, r_par, r_trans, z, num_pairs


def compute_xi_batch(healpixs, normalise=True, to_host=True):
    """One launch for many HEALPix pixels: row k equals ``compute_xi([healpixs[k]])``."""
    healpixs = list(healpixs)
    eng, host1, dev1, host2, dev2 = _catalogs()
    params = params_from_module(_THIS, cross=True)
    pairs = _pairs_for(healpixs)
    rows = np.concatenate([np.full(host1.first_of(hp)[1] - host1.first_of(hp)[0], k, np.int32)
                           for k, hp in enumerate(healpixs)]) if healpixs else np.zeros(0, np.int32)
    out = eng.xi(dev1, dev2, params, pairs, rows, len(healpixs), cross_obj=True,
                 variant=_XI_VARIANT, normalise=normalise)
    host = out.cpu().numpy() if to_host else out  # to_host=False: device tensor (multi-GPU gather)
    _corr.bump_progress(_THIS, pairs.n_f1, userprint)
    _corr.clear_neighbours(host1.objs, pairs.f1_index.cpu().numpy())
    _STORE.drop(healpixs)
    return host


def _single_forest_and_objects(z1, r_comov1, dist_m1, weights1, delta1, log_lambda1, order1, z2,
                               r_comov2, dist_m2, weights2):
    d1 = _Delta(1, 0., 0., 0., 1, 0, 1, np.asarray(log_lambda1, dtype=np.float64),
                np.asarray(weights1, dtype=np.float64), np.asarray(delta1, dtype=np.float64),
                order1)
    d1.z, d1.r_comov, d1.dist_m = (np.asarray(z1, dtype=np.float64),
                                   np.asarray(r_comov1, dtype=np.float64),
                                   np.asarray(dist_m1, dtype=np.float64))
    qs = []
    for k in range(len(z2)):
        q = _QSO(100 + k, 0., 0., float(z2[k]), 2, 0, 2)
        q.weights, q.r_comov, q.dist_m = float(weights2[k]), float(r_comov2[k]), float(dist_m2[k])
        qs.append(q)
    return d1, qs


def _explicit_pairs(eng, ang):
    from .engine import PairList
    torch = eng.torch
    ang = np.ascontiguousarray(ang, dtype=np.float64)
    m = ang.size
    i32 = lambda v: torch.as_tensor(np.asarray(v, dtype=np.int32), device=eng.device)
    f64 = lambda v: torch.as_tensor(np.asarray(v, dtype=np.float64), device=eng.device)
    return PairList(eng, i32([0]), torch.tensor([0, m], dtype=torch.int64, device=eng.device),
                    i32(np.zeros(m)), i32(np.arange(m)), f64(ang), f64(np.cos(ang / 2)),
                    f64(np.sin(ang / 2)))


def compute_xi_forest_pairs_fast(z1, r_comov1, dist_m1, weights1, delta1, z2, r_comov2, dist_m2,
                                 weights2, ang, rebin_weight, rebin_xi, rebin_r_par, rebin_r_trans,
                                 rebin_z, rebin_num_pairs):
    """One forest against a list of objects, accumulated in place (xcf.py:223-322).  Kept for
    signature parity; compute_xi does not go through it."""
    eng = get_engine()
    d1, qs = _single_forest_and_objects(z1, r_comov1, dist_m1, weights1, delta1,
                                        np.zeros(len(z1)), 0, z2, r_comov2, dist_m2, weights2)
    dev1 = eng.device_catalog(_catalog.pack({0: [d1]}), cache=False)
    dev2 = eng.device_catalog(_catalog.pack({0: qs}, is_object=True), cache=False)
    params = params_from_module(_THIS, cross=True)
    pairs = _explicit_pairs(eng, ang)
    row = eng.torch.zeros(1, dtype=eng.torch.int32, device=eng.device)
    out = eng.xi(dev1, dev2, params, pairs, row, 1, cross_obj=True,
                 variant=_XI_VARIANT).cpu().numpy()[0]
    rebin_weight += out[0]
    rebin_xi += out[1]
    rebin_r_par += out[2]
    rebin_r_trans += out[3]
    rebin_z += out[4]
    rebin_num_pairs += out[5].view(np.int64)


compute_xi_forest_pairs = compute_xi_forest_pairs_fast


def compute_dmat_forest_pairs_fast(log_lambda1, r_comov1, dist_m1, z1, weights1, r_comov2, dist_m2,
                                   z2, weights2, ang, weights_dmat, dmat, r_par_eff, r_trans_eff,
                                   z_eff, weight_eff, order1):
    """One forest against its kept objects, accumulated in place (xcf.py:427-674; ``dmat`` is the
    flat [nb * nbm] array).  Kept for signature parity; compute_dmat does not go through it."""
    eng = get_engine()
    d1, qs = _single_forest_and_objects(z1, r_comov1, dist_m1, weights1, np.zeros(len(z1)),
                                        log_lambda1, int(order1), z2, r_comov2, dist_m2, weights2)
    dev1 = eng.device_catalog(_catalog.pack({0: [d1]}), cache=False)
    dev2 = eng.device_catalog(_catalog.pack({0: qs}, is_object=True), cache=False)
    params = params_from_module(_THIS, cross=True)
    pairs = _explicit_pairs(eng, ang)
    res = [t.cpu().numpy() for t in eng.dmat(dev1, dev2, params, pairs, cross_obj=True)]
    weights_dmat += res[0]
    dmat += res[1].reshape(dmat.shape)
    r_par_eff += res[2]
    r_trans_eff += res[3]
    z_eff += res[4]
    weight_eff += res[5]


compute_dmat_forest_pairs = compute_dmat_forest_pairs_fast


def compute_dmat(healpixs):
    """Distortion matrix of the cross-correlation (xcf.py:325-424).  The --rej draw uses the global
    legacy NumPy RNG in the reference's order (xcf.py:379); forests whose draw keeps nothing are
    not counted (xcf.py:380-383).  Returns the reference's 8-tuple."""
    healpixs = list(healpixs)
    eng, host1, dev1, host2, dev2 = _catalogs()
    params = params_from_module(_THIS, cross=True)
    pairs = _pairs_for(healpixs)
    f1_index = pairs.f1_index.cpu().numpy()
    if np.any(host1.arrays["order"][f1_index] < 0):
        raise RuntimeError("Trying to compute the distortion matrix but "
                           "order is not defined for the deltas. "
                           "Check previous warning to solve this issue")  # xcf.py:368-373
    offset = pairs.host_offset()
    keep = np.random.rand(int(offset[-1])) > reject  # xcf.py:379, one stream for all forests
    kept_per_forest = np.add.reduceat(np.append(keep.astype(np.int64), 0), offset[:-1]) \
        if len(offset) > 1 else np.zeros(0, dtype=np.int64)
    kept_per_forest = np.where(np.diff(offset) > 0, kept_per_forest, 0)
    counted = kept_per_forest > 0  # xcf.py:380-383
    num_pairs = int(np.diff(offset)[counted].sum())
    num_pairs_used = int(kept_per_forest.sum())
    pairs.nb_keep = eng.torch.from_numpy(keep.astype(np.uint8)).to(eng.device)
    res = eng.dmat(dev1, dev2, params, pairs, cross_obj=True)
    weights_dmat, dmat, r_par_eff, r_trans_eff, z_eff, weight_eff = (t.cpu().numpy() for t in res)
    _corr.bump_progress(_THIS, pairs.n_f1, userprint)
    for k, f1 in enumerate(f1_index):
        if counted[k]:
            _corr.set_neighbours(host1.objs[f1], None)  # xcf.py:409 (skipped forests keep theirs)
    _STORE.drop(healpixs)
    return (weights_dmat, dmat, r_par_eff, r_trans_eff, z_eff, weight_eff, num_pairs,
            num_pairs_used)


def compute_wick_terms(healpixs):
    """Wick expansion of the covariance matrix of the cross-correlation, diagrams T1-T4
    (xcf.py:838-941; compute_wickT1234_pairs :1219-1351).  The per-forest --rej draw uses the
    global legacy NumPy RNG in the reference's order (xcf.py:888-890).  Returns (weights_wick,
    num_pairs_wick, num_pairs, num_pairs_used, t1..t6); t5, t6 (``xi_wick`` given and
    ``max_diagram > 4``: four-point diagrams over pairs of forests) are not built --
    NotImplementedError."""
    import ctypes
    from . import _lib, _wick
    healpixs = list(healpixs)
    if xi_wick is not None and max_diagram is not None and max_diagram > 4:
        raise NotImplementedError("picca_b200: Wick diagrams T5-T6 (max_diagram > 4) are not "
                                  "implemented on the B200 path")
    eng, host1, dev1, host2, dev2 = _catalogs()
    params = params_from_module(_THIS, cross=True)
    pairs = _pairs_for(healpixs)
    torch = eng.torch
    nb = params.num_bins_r_par * params.num_bins_r_trans
    num_pairs = 0
    num_pairs_used = 0
    keep_forest = []
    for healpix in healpixs:
        num_pairs += len(data[healpix])
        w = np.random.rand(len(data[healpix])) > reject   # xcf.py:889
        num_pairs_used += int(w.sum())
        keep_forest.append(w)
    keep_forest = np.concatenate(keep_forest) if keep_forest else np.zeros(0, dtype=bool)
    zeros = lambda *shape: torch.zeros(shape, dtype=torch.float64, device=eng.device)
    t1, t2, t3, t4 = zeros(nb, nb), zeros(nb, nb), zeros(nb, nb), zeros(nb, nb)
    weights_wick = zeros(nb)
    num_pairs_wick = torch.zeros(nb, dtype=torch.int64, device=eng.device)
    counts = np.diff(pairs.host_offset())
    if pairs.n_pairs and keep_forest.any() and counts[keep_forest].max(initial=0) > 0:
        var1, ze1, xb1, xy1, n_x1 = _wick.pixel_inputs(eng, host1, get_variance_1d, xi_1d, z_ref,
                                                       alpha)
        ze_obj = torch.from_numpy(np.ascontiguousarray(
            ((1 + host2.arrays["z_qso"]) / (1 + z_ref))**(alpha_obj - 1))).to(eng.device)
        keep_dev = torch.from_numpy(keep_forest.astype(np.uint8)).to(eng.device)
        max_nb = int(counts[keep_forest].max())
        nbytes = int(eng.lib.pb2_wick_scratch_bytes(ctypes.c_int64(host1.max_pix),
                                                    ctypes.c_int64(max_nb), ctypes.c_int32(1)))
        scratch = torch.empty(nbytes, dtype=torch.uint8, device=eng.device)
        ps = pairs.struct()
        ptr = lambda t: ctypes.c_void_p(t.data_ptr())
        _lib.check(eng.lib.pb2_wick_cross(
            ctypes.byref(dev1.struct), ctypes.byref(dev2.struct), ctypes.byref(params),
            ctypes.byref(ps), ptr(keep_dev), ctypes.c_int64(max_nb), ptr(var1), ptr(ze1),
            ptr(ze_obj), ctypes.c_int32(n_x1), ptr(xb1), ptr(xy1), ptr(weights_wick),
            ptr(num_pairs_wick), ptr(t1), ptr(t2), ptr(t3), ptr(t4), ptr(scratch),
            ctypes.c_int64(nbytes), eng.stream_ptr()), "pb2_wick_cross")
    _corr.bump_progress(_THIS, int(keep_forest.sum()), userprint)
    _STORE.drop(healpixs)
    host_t = [t.cpu().numpy() for t in (t1, t2, t3, t4)]
    if np.isnan(np.diagonal(host_t[0])).any():
        # weights12**2 / weight1 with a zero-weight pixel in range (xcf.py:1316): Numba raises
        raise ZeroDivisionError("division by zero")
    empty = lambda: np.zeros((nb, nb))
    return (weights_wick.cpu().numpy(), num_pairs_wick.cpu().numpy(), num_pairs, num_pairs_used,
            host_t[0], host_t[1], host_t[2], host_t[3], empty(), empty())


def compute_metal_dmat(healpixs, abs_igm="SiII(1526)"):
    """Metal distortion matrix of the cross-correlation (xcf.py:677-835): data bins from the
    Lyman-alpha distances of the forest pixels, model bins from the distances they would have if
    the absorption came from ``abs_igm``.  Forests without a pixel consistent with the quasar
    redshift are skipped BEFORE the --rej draw (xcf.py:741-742): they consume no random numbers
    and are not counted.  Returns the reference's 8-tuple of un-normalised sums."""
    import ctypes
    from . import _lib
    from .cf import _absorber_wavelength
    healpixs = list(healpixs)
    eng, host1, dev1, host2, dev2 = _catalogs()
    params = params_from_module(_THIS, cross=True)
    pairs = _pairs_for(healpixs)
    f1_index = pairs.f1_index.cpu().numpy()
    offset = pairs.host_offset()
    # per pixel: xcf.py:729-731 and :769-771, evaluated on the host like the reference
    z_abs = 10**host1.arrays["log_lambda"] / _absorber_wavelength(abs_igm) - 1
    r_comov_abs = np.asarray(cosmo.get_r_comov(z_abs), dtype=np.float64)
    dist_m_abs = np.asarray(cosmo.get_dist_m(z_abs), dtype=np.float64)
    evol = ((1.0 + z_abs) / (1.0 + z_ref))**(alpha_abs[abs_igm] - 1.0)
    # forests with at least one pixel with z_abs < z_qso (xcf.py:735-742)
    pix_off = host1.arrays["offset"]
    valid_pix = z_abs < np.repeat(host1.arrays["z_qso"], np.diff(pix_off))
    n_valid = np.add.reduceat(np.append(valid_pix.astype(np.int64), 0), pix_off[:-1])
    n_valid = np.where(np.diff(pix_off) > 0, n_valid, 0)
    drawn = n_valid[f1_index] > 0
    lengths = np.diff(offset)
    keep = np.zeros(int(offset[-1]), dtype=bool)
    draw = np.random.rand(int(lengths[drawn].sum())) > reject  # xcf.py:744, drawn forests only
    pair_drawn = np.repeat(drawn, lengths)
    keep[pair_drawn] = draw
    num_pairs = int(lengths[drawn].sum())
    num_pairs_used = int(keep.sum())
    pairs.nb_keep = eng.torch.from_numpy(keep.astype(np.uint8)).to(eng.device)

    torch = eng.torch
    nb = params.num_bins_r_par * params.num_bins_r_trans
    nbm = params.num_model_bins_r_par * params.num_model_bins_r_trans
    zeros = lambda *shape: torch.zeros(shape, dtype=torch.float64, device=eng.device)
    weights_dmat, dmat = zeros(nb), zeros(nb, nbm)
    r_par_eff, r_trans_eff, z_eff, weight_eff = zeros(nbm), zeros(nbm), zeros(nbm), zeros(nbm)
    up = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(eng.device)
    m = [up(a) for a in (z_abs, r_comov_abs, dist_m_abs, evol)]
    ps = pairs.struct()
    ptr = lambda t: ctypes.c_void_p(t.data_ptr())
    _lib.check(eng.lib.pb2_metal_dmat_cross(
        ctypes.byref(dev1.struct), ctypes.byref(dev2.struct), ctypes.byref(params),
        ctypes.byref(ps), ptr(m[0]), ptr(m[1]), ptr(m[2]), ptr(m[3]), ptr(weights_dmat),
        ptr(dmat), ptr(r_par_eff), ptr(r_trans_eff), ptr(z_eff), ptr(weight_eff),
        eng.stream_ptr()), "pb2_metal_dmat_cross")
    res = tuple(t.cpu().numpy() for t in (weights_dmat, dmat, r_par_eff, r_trans_eff, z_eff,
                                          weight_eff))
    _corr.bump_progress(_THIS, pairs.n_f1, userprint)
    for k, f1 in enumerate(f1_index):
        if drawn[k]:
            _corr.set_neighbours(host1.objs[f1], None)  # xcf.py:819 (skipped forests keep theirs)
    _STORE.drop(healpixs)
    return res + (num_pairs, num_pairs_used)
