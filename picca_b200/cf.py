"""B200 implementation of the hot path of ``picca.cf`` behind the reference's module API.

Same module globals (reference py/picca/cf.py:28-79), same functions, same return tuples:

    fill_neighs(healpixs)                         cf.py:82-135
    compute_xi(healpixs) -> 6-tuple               cf.py:138-247
    compute_dmat(healpixs) -> 8-tuple             cf.py:390-517
    compute_xi_forest_pairs_fast(...)             cf.py:250-387  (in-place accumulate)
    compute_dmat_forest_pairs_fast(...)           cf.py:520-887  (in-place accumulate)
    compute_metal_dmat(healpixs, abs_igm1, abs_igm2) -> 8-tuple   cf.py:890-1232

so that picca_cf.py / picca_dmat.py run unchanged once this module is importable as ``picca.cf``
(see ``picca_b200.overlay`` and INTEGRATION.md).  All arithmetic runs in the CUDA kernels of
libpicca_b200.so; there is no Numba and no CPU fallback.  Globals are read at call time.
"""
import sys

import numpy as np

from . import _corr, catalog as _catalog
from .engine import MODE_AUTO, MODE_CROSS, get_engine
from .forest import Delta as _Delta
from .params import params_from_module


def userprint(*args, **kwds):
    """reference py/picca/utils.py:20-28"""
    print(*args, **kwds)
    sys.stdout.flush()


# ---- module globals: names and defaults of reference cf.py:28-79
num_bins_r_par = None
num_bins_r_trans = None
num_model_bins_r_trans = None
num_model_bins_r_par = None
r_par_max = None
r_par_min = None
z_min_pairs = None
z_max_pairs = None
r_trans_max = None
ang_max = None
nside = None

zerr_cut_deg = None
zerr_cut_kms = None

counter = None
num_data = None
num_data2 = None

z_ref = None
alpha = None
alpha2 = None
alpha_abs = None
lambda_abs = None
lambda_abs2 = None

data = None
data2 = None

cosmo = None

reject = None
lock = None
x_correlation = False
rmu_binning = False
ang_correlation = False
remove_same_half_plate_close_pairs = False

# variables for distortion matrix
redshift_evolution_in_distortion_matrix = True

# variables used in the 1D correlation function analysis (kept for attribute parity)
num_pixels = None
log_lambda_min = None
log_lambda_max = None
delta_log_lambda = None

# variables used in the wick covariance matrix computation (kept for attribute parity)
get_variance_1d = {}
xi_1d = {}
max_diagram = None
xi_wick = {}

# rest wavelengths (Angstrom) of the transitions compute_metal_dmat is usually run with; the
# reference's full table (``picca.constants.ABSORBER_IGM``) is used when it is importable, and
# entries can be added here (same names)
absorber_igm = {"LYA": 1215.67, "LYB": 1025.72, "SiIII(1207)": 1206.500,
                "SiII(1190)": 1190.4158, "SiII(1193)": 1193.2897, "SiII(1260)": 1260.4221,
                "CIV(eff)": 1549.06}

_THIS = sys.modules[__name__]
_STORE = _corr.NeighbourStore()
# xi kernel variant: 0 = product (diagonal sweep), 1 = brute-force validation kernel
_XI_VARIANT = 0


def _catalogs():
    eng, host1, dev1 = _corr.engine_and_catalog(data, ang_correlation=ang_correlation)
    if data2 is not None:
        _, host2, dev2 = _corr.engine_and_catalog(data2, ang_correlation=ang_correlation)
    else:
        host2, dev2 = host1, dev1
    return eng, host1, dev1, host2, dev2


def _check_half_plate(host1, host2):
    if remove_same_half_plate_close_pairs and not (host1.ids_are_int and host2.ids_are_int):
        raise RuntimeError("Trying to figure out if two spectra "
                           "come from the same half plate but "
                           "combined reobservations were given")  # cf.py:172-179


def fill_neighs(healpixs):
    """Create the neighbour list of every forest of ``healpixs`` (cf.py:82-135) on the device."""
    healpixs = list(healpixs)
    if _corr.defer_fill():   # main process, no CUDA yet: see _corr.defer_fill
        _STORE.defer(healpixs, data, _fill_neighs_now)
        return
    _fill_neighs_now(healpixs)


def _fill_neighs_now(healpixs):
    healpixs = list(healpixs)
    eng, host1, dev1, host2, dev2 = _catalogs()
    params = params_from_module(_THIS)
    index, ranges = _corr.forest_index_of(host1, healpixs)
    mode = MODE_CROSS if data2 is not None else MODE_AUTO
    pairs = eng.neighbours(dev1, dev2, params, mode, index)
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
    """Correlation function of the forests of ``healpixs`` with their neighbours (cf.py:138-247).

    Returns (weights, xi, r_par, r_trans, z, num_pairs), normalised per call like the reference.
    """
    healpixs = list(healpixs)
    eng, host1, dev1, host2, dev2 = _catalogs()
    _check_half_plate(host1, host2)
    params = params_from_module(_THIS)
    pairs = _pairs_for(healpixs)
    out_row = eng.torch.zeros(pairs.n_f1, dtype=eng.torch.int32, device=eng.device)
    out = eng.xi(dev1, dev2, params, pairs, out_row, 1, variant=_XI_VARIANT, normalise=True)
    host = out.cpu().numpy()[0]
    _corr.bump_progress(_THIS, pairs.n_f1, userprint)
    _corr.clear_neighbours(host1.objs, pairs.f1_index.cpu().numpy())  # cf.py:240
    _STORE.drop(healpixs)
    weights, xi, r_par, r_trans, z = (np.ascontiguousarray(host[k]) for k in range(5))
    num_pairs = np.ascontiguousarray(host[5]).view(np.int64)
    return weights, xi, r_par, r_trans, z, num_pairs


def compute_xi_batch(healpixs, normalise=True, to_host=True):
    """One launch for many HEALPix pixels: row k of the result is ``compute_xi([healpixs[k]])``.
    Returns an array [len(healpixs), 6, nb] (row 5 = int64 counts viewed as float64 slots) --
    what picca_cf.py stacks from its Pool.map (picca_cf.py:466-473)."""
    healpixs = list(healpixs)
    eng, host1, dev1, host2, dev2 = _catalogs()
    _check_half_plate(host1, host2)
    params = params_from_module(_THIS)
    pairs = _pairs_for(healpixs)
    rows = np.concatenate([np.full(host1.first_of(hp)[1] - host1.first_of(hp)[0], k, np.int32)
                           for k, hp in enumerate(healpixs)]) if healpixs else np.zeros(0, np.int32)
    out = eng.xi(dev1, dev2, params, pairs, rows, len(healpixs), variant=_XI_VARIANT,
                 normalise=normalise)
    host = out.cpu().numpy() if to_host else out  # to_host=False: device tensor (multi-GPU gather)
    _corr.bump_progress(_THIS, pairs.n_f1, userprint)
    _corr.clear_neighbours(host1.objs, pairs.f1_index.cpu().numpy())
    _STORE.drop(healpixs)
    return host


def compute_xi_forest_pairs_fast(z1, r_comov1, dist_m1, weights1, delta1, z_qso_1, z2, r_comov2,
                                 dist_m2, weights2, delta2, z_qso_2, ang, same_half_plate,
                                 rebin_weight, rebin_xi, rebin_r_par, rebin_r_trans, rebin_z,
                                 rebin_num_pairs):
    """One forest pair, accumulated in place into the caller's rebin arrays (cf.py:250-387).
    Kept for signature parity; compute_xi does not go through it."""
    from .engine import PairList
    eng = get_engine()
    torch = eng.torch

    def one(z, rc, dm, w, de, zq, plate):
        d = _Delta(1, 0., 0., zq, plate, 0, 1 if same_half_plate else plate, np.zeros(len(z)),
                   np.asarray(w, dtype=np.float64), np.asarray(de, dtype=np.float64), 0)
        d.z, d.r_comov, d.dist_m = (np.asarray(z, dtype=np.float64),
                                    np.asarray(rc, dtype=np.float64),
                                    np.asarray(dm, dtype=np.float64))
        return d

    # plate/fiberid chosen so that pb2_same_half_plate() reproduces the caller's flag
    d1 = one(z1, r_comov1, dist_m1, weights1, delta1, z_qso_1, 1)
    d2 = one(z2, r_comov2, dist_m2, weights2, delta2, z_qso_2, 1 if same_half_plate else 2)
    dev1 = eng.device_catalog(_catalog.pack({0: [d1]}), cache=False)
    dev2 = eng.device_catalog(_catalog.pack({0: [d2]}), cache=False)
    params = params_from_module(_THIS)
    i32 = lambda v: torch.tensor(v, dtype=torch.int32, device=eng.device)
    f64 = lambda v: torch.tensor(v, dtype=torch.float64, device=eng.device)
    ang = float(ang)
    pairs = PairList(eng, i32([0]), torch.tensor([0, 1], dtype=torch.int64, device=eng.device),
                     i32([0]), i32([0]), f64([ang]), f64([np.cos(ang / 2)]),
                     f64([np.sin(ang / 2)]))
    out = eng.xi(dev1, dev2, params, pairs, i32([0]), 1, variant=_XI_VARIANT).cpu().numpy()[0]
    rebin_weight += out[0]
    rebin_xi += out[1]
    rebin_r_par += out[2]
    rebin_r_trans += out[3]
    rebin_z += out[4]
    rebin_num_pairs += out[5].view(np.int64)


# older picca spelling (module docstring of the reference, cf.py:5-9)
compute_xi_forest_pairs = compute_xi_forest_pairs_fast


def compute_dmat_forest_pairs_fast(log_lambda1, log_lambda2, r_comov1, r_comov2, dist_m1, dist_m2,
                                   z1, z2, weights1, weights2, z_qso_1, z_qso_2, ang, weights_dmat,
                                   dmat, r_par_eff, r_trans_eff, z_eff, weight_eff,
                                   same_half_plate, order1, order2):
    """One forest pair of the distortion matrix, accumulated in place into the caller's arrays
    (cf.py:520-887; ``dmat`` is the flat [nb * nbm] array of cf.py:410-415).  Kept for signature
    parity; compute_dmat does not go through it."""
    from .engine import PairList
    eng = get_engine()
    torch = eng.torch

    def one(ll, z, rc, dm, w, zq, plate, order):
        d = _Delta(1, 0., 0., zq, plate, 0, 1, np.asarray(ll, dtype=np.float64),
                   np.asarray(w, dtype=np.float64), np.zeros(len(z)), int(order))
        d.z, d.r_comov, d.dist_m = (np.asarray(z, dtype=np.float64),
                                    np.asarray(rc, dtype=np.float64),
                                    np.asarray(dm, dtype=np.float64))
        return d

    d1 = one(log_lambda1, z1, r_comov1, dist_m1, weights1, z_qso_1, 1, order1)
    d2 = one(log_lambda2, z2, r_comov2, dist_m2, weights2, z_qso_2, 1 if same_half_plate else 2,
             order2)
    dev1 = eng.device_catalog(_catalog.pack({0: [d1]}), cache=False)
    dev2 = eng.device_catalog(_catalog.pack({0: [d2]}), cache=False)
    params = params_from_module(_THIS)
    i32 = lambda v: torch.tensor(v, dtype=torch.int32, device=eng.device)
    f64 = lambda v: torch.tensor(v, dtype=torch.float64, device=eng.device)
    ang = float(ang)
    pairs = PairList(eng, i32([0]), torch.tensor([0, 1], dtype=torch.int64, device=eng.device),
                     i32([0]), i32([0]), f64([ang]), f64([np.cos(ang / 2)]),
                     f64([np.sin(ang / 2)]))
    res = [t.cpu().numpy() for t in eng.dmat(dev1, dev2, params, pairs)]
    weights_dmat += res[0]
    dmat += res[1].reshape(dmat.shape)
    r_par_eff += res[2]
    r_trans_eff += res[3]
    z_eff += res[4]
    weight_eff += res[5]


compute_dmat_forest_pairs = compute_dmat_forest_pairs_fast


def compute_dmat(healpixs):
    """Distortion matrix of the forests of ``healpixs`` (cf.py:390-517).  The --rej draw uses the
    global legacy NumPy RNG in the reference's order (cf.py:444), so results match the reference
    for the same seed and chunking.  Returns the reference's 8-tuple of un-normalised sums."""
    healpixs = list(healpixs)
    eng, host1, dev1, host2, dev2 = _catalogs()
    _check_half_plate(host1, host2)
    params = params_from_module(_THIS)
    pairs = _pairs_for(healpixs)
    f1_index = pairs.f1_index.cpu().numpy()
    order1 = host1.arrays["order"]
    if np.any(order1[f1_index] < 0):
        raise RuntimeError("Trying to compute the distortion matrix but "
                           "order is not defined for the deltas. "
                           "Check previous warning to solve this issue")  # cf.py:433-438
    offset = pairs.host_offset()
    # one draw per forest, in catalogue order, len(neighbours) numbers each (cf.py:444);
    # consecutive rand(n) calls consume the MT19937 stream exactly like one rand(sum n)
    keep = np.random.rand(int(offset[-1])) > reject
    num_pairs = int(offset[-1])
    num_pairs_used = int(keep.sum())
    if np.any(host2.arrays["order"][pairs.host_f2()[keep]] < 0):
        raise RuntimeError("Trying to compute the distortion matrix but "
                           "order is not defined for the deltas. "
                           "Check previous warning to solve this issue")  # cf.py:464-469
    pairs.nb_keep = eng.torch.from_numpy(keep.astype(np.uint8)).to(eng.device)
    res = eng.dmat(dev1, dev2, params, pairs)
    weights_dmat, dmat, r_par_eff, r_trans_eff, z_eff, weight_eff = (t.cpu().numpy() for t in res)
    _corr.bump_progress(_THIS, pairs.n_f1, userprint)
    _corr.clear_neighbours(host1.objs, f1_index)  # cf.py:502
    _STORE.drop(healpixs)
    return (weights_dmat, dmat, r_par_eff, r_trans_eff, z_eff, weight_eff, num_pairs,
            num_pairs_used)


def compute_wick_terms(healpixs):
    """Wick expansion of the covariance matrix, diagrams T1-T3 (cf.py:1326-1494 with
    ``max_diagram <= 3``; compute_wickT123_pairs :1497-1626).  The per-forest --rej draw uses the
    global legacy NumPy RNG in the reference's order (cf.py:1378: one number per forest of each
    HEALPix pixel).  Returns (weights_wick, num_pairs_wick, num_pairs, num_pairs_used, t1..t6);
    t4-t6 need ``max_diagram > 3`` (three-forest diagrams): not built -- NotImplementedError."""
    import ctypes
    from . import _lib, _wick
    healpixs = list(healpixs)
    if max_diagram is not None and max_diagram > 3:
        raise NotImplementedError("picca_b200: Wick diagrams T4-T6 (max_diagram > 3) are not "
                                  "implemented on the B200 path")
    eng, host1, dev1, host2, dev2 = _catalogs()
    params = params_from_module(_THIS)
    pairs = _pairs_for(healpixs)
    torch = eng.torch
    nb = params.num_bins_r_par * params.num_bins_r_trans
    num_pairs = 0
    num_pairs_used = 0
    keep_forest = []
    for healpix in healpixs:
        w = np.random.rand(len(data[healpix])) > reject   # cf.py:1378
        num_pairs += len(data[healpix])
        num_pairs_used += int(w.sum())
        keep_forest.append(w)
    keep_forest = np.concatenate(keep_forest) if keep_forest else np.zeros(0, dtype=bool)
    keep_dev = torch.from_numpy(keep_forest.astype(np.uint8)).to(eng.device)
    pairs.nb_keep = keep_dev[pairs.nb_f1.to(torch.int64)].contiguous()   # pair kept iff its forest is
    zeros = lambda *shape: torch.zeros(shape, dtype=torch.float64, device=eng.device)
    t1, t2, t3 = zeros(nb, nb), zeros(nb, nb), zeros(nb, nb)
    weights_wick = zeros(nb)
    num_pairs_wick = torch.zeros(nb, dtype=torch.int64, device=eng.device)
    if pairs.n_pairs and keep_forest.any():
        in1 = _wick.pixel_inputs(eng, host1, get_variance_1d, xi_1d, z_ref, alpha)
        in2 = in1 if host2 is host1 else _wick.pixel_inputs(eng, host2, get_variance_1d, xi_1d,
                                                            z_ref, alpha2)
        if host2 is host1 and alpha2 is not None and alpha2 != alpha:  # cf.py:1557 uses alpha2
            ze2 = torch.from_numpy(np.ascontiguousarray(
                ((1 + host2.arrays["z"]) / (1 + z_ref))**(alpha2 - 1))).to(eng.device)
            in2 = (in1[0], ze2) + in1[2:]
        nbytes = int(eng.lib.pb2_wick_scratch_bytes(ctypes.c_int64(host1.max_pix),
                                                    ctypes.c_int64(host2.max_pix),
                                                    ctypes.c_int32(0)))
        scratch = torch.empty(nbytes, dtype=torch.uint8, device=eng.device)
        ps = pairs.struct()
        ptr = lambda t: ctypes.c_void_p(t.data_ptr())
        _lib.check(eng.lib.pb2_wick_auto(
            ctypes.byref(dev1.struct), ctypes.byref(dev2.struct), ctypes.byref(params),
            ctypes.byref(ps), ptr(in1[0]), ptr(in1[1]), ptr(in2[0]), ptr(in2[1]),
            ctypes.c_int32(in1[4]), ptr(in1[2]), ptr(in1[3]), ctypes.c_int32(in2[4]), ptr(in2[2]),
            ptr(in2[3]), ptr(weights_wick), ptr(num_pairs_wick), ptr(t1), ptr(t2), ptr(t3),
            ptr(scratch), ctypes.c_int64(nbytes), eng.stream_ptr()), "pb2_wick_auto")
    _corr.bump_progress(_THIS, int(keep_forest.sum()), userprint)
    _STORE.drop(healpixs)
    host_t = [t.cpu().numpy() for t in (t1, t2, t3)]
    empty = lambda: np.zeros((nb, nb))
    return (weights_wick.cpu().numpy(), num_pairs_wick.cpu().numpy(), num_pairs, num_pairs_used,
            host_t[0], host_t[1], host_t[2], empty(), empty(), empty())


def _absorber_wavelength(name):
    if name in absorber_igm:
        return absorber_igm[name]
    try:
        from picca import constants  # the reference's table, when installed / overlaid
        return constants.ABSORBER_IGM[name]
    except ImportError:
        raise KeyError(name)


def _metal_arrays(eng, host, name):
    """Per pixel of a packed catalogue, for absorber ``name``: the redshift the pixel would have
    (cf.py:944), its distances on the fiducial cosmology (cf.py:945-946) and the evolution factor
    (1+z)**(alpha_abs-1) (cf.py:1033-1035), evaluated on the host with NumPy / the caller's
    cosmology exactly as the reference does, then placed in HBM."""
    z_abs = 10**host.arrays["log_lambda"] / _absorber_wavelength(name) - 1
    arrays = (z_abs, np.asarray(cosmo.get_r_comov(z_abs), dtype=np.float64),
              np.asarray(cosmo.get_dist_m(z_abs), dtype=np.float64),
              (1 + z_abs)**(alpha_abs[name] - 1))
    return tuple(eng.torch.from_numpy(np.ascontiguousarray(a)).to(eng.device) for a in arrays)


def compute_metal_dmat(healpixs, abs_igm1="LYA", abs_igm2="SiIII(1207)"):
    """Metal distortion matrix of the forests of ``healpixs`` (cf.py:890-1232): data bins from
    the Lyman-alpha distances, model bins from the distances of the (abs_igm1, abs_igm2)
    absorptions.  Same --rej draw, same 8-tuple of un-normalised sums as the reference."""
    import ctypes
    from . import _lib
    healpixs = list(healpixs)
    eng, host1, dev1, host2, dev2 = _catalogs()
    _check_half_plate(host1, host2)
    params = params_from_module(_THIS)
    pairs = _pairs_for(healpixs)
    f1_index = pairs.f1_index.cpu().numpy()
    offset = pairs.host_offset()
    keep = np.random.rand(int(offset[-1])) > reject  # cf.py:944, one draw per forest in order
    num_pairs = int(offset[-1])
    num_pairs_used = int(keep.sum())
    pairs.nb_keep = eng.torch.from_numpy(keep.astype(np.uint8)).to(eng.device)

    torch = eng.torch
    nb = params.num_bins_r_par * params.num_bins_r_trans
    nbm = params.num_model_bins_r_par * params.num_model_bins_r_trans
    zeros = lambda *shape: torch.zeros(shape, dtype=torch.float64, device=eng.device)
    weights_dmat, dmat = zeros(nb), zeros(nb, nbm)
    r_par_eff, r_trans_eff, z_eff, weight_eff = zeros(nbm), zeros(nbm), zeros(nbm), zeros(nbm)
    cache = {}

    def arrays(host, name):
        key = (id(host), name)
        if key not in cache:
            cache[key] = _metal_arrays(eng, host, name)
        return cache[key]

    passes = [(abs_igm1, abs_igm2)]
    if ((not x_correlation) and (abs_igm1 != abs_igm2)) or \
            (x_correlation and (lambda_abs == lambda_abs2)):  # cf.py:1089-1091
        passes.append((abs_igm2, abs_igm1))
    ps = pairs.struct()
    ptr = lambda t: ctypes.c_void_p(t.data_ptr())
    for name1, name2 in passes:
        m1, m2 = arrays(host1, name1), arrays(host2, name2)
        den = (1 + z_ref)**(alpha_abs[abs_igm1] + alpha_abs[abs_igm2] - 2)  # cf.py:1036, :1168
        _lib.check(eng.lib.pb2_metal_dmat_auto(
            ctypes.byref(dev1.struct), ctypes.byref(dev2.struct), ctypes.byref(params),
            ctypes.byref(ps), ptr(m1[0]), ptr(m1[1]), ptr(m1[2]), ptr(m1[3]), ptr(m2[0]),
            ptr(m2[1]), ptr(m2[2]), ptr(m2[3]), ctypes.c_double(den), ptr(weights_dmat),
            ptr(dmat), ptr(r_par_eff), ptr(r_trans_eff), ptr(z_eff), ptr(weight_eff),
            eng.stream_ptr()), "pb2_metal_dmat_auto")
    res = tuple(t.cpu().numpy() for t in (weights_dmat, dmat, r_par_eff, r_trans_eff, z_eff,
                                          weight_eff))
    _corr.bump_progress(_THIS, pairs.n_f1, userprint)
    _corr.clear_neighbours(host1.objs, f1_index)  # cf.py:1219
    _STORE.drop(healpixs)
    return res + (num_pairs, num_pairs_used)
