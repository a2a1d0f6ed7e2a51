"""Deterministic synthetic forests / quasars of the shapes BASELINE.json names.

The generator produces exactly what ``picca.io.read_deltas`` / ``read_objects`` hand to the hot
path (reference py/picca/io.py:485-512, :595-612): ``dict[healpix] -> list[Delta]`` with z, r_comov,
dist_m filled, weights scaled by ((1+z)/(1+z_ref))**(alpha-1) and deltas projected.  It does not
read or need the reference.
"""
import numpy as np

from .forest import PIXEL_FIELDS, Delta, QSO, register_soa

LYA = 1215.67  # reference py/picca/constants.py ABSORBER_IGM["LYA"]
SPEED_LIGHT = 299792.458


class FlatLCDM:
    """Tabulated flat-LCDM distances in Mpc/h (same construction as reference
    py/picca/constants.py:193-229: 10000-point trapezoid table to z=10, linear interpolation)."""

    def __init__(self, Om=0.315):
        num_bins, z_max = 10000, 10.
        dz = z_max / num_bins
        z = np.arange(num_bins, dtype=float) * dz
        hubble = 100. * np.sqrt((1. - Om) + Om * (1. + z)**3)
        r = np.zeros(num_bins)
        r[1:] = np.cumsum(SPEED_LIGHT * (1. / hubble[:-1] + 1. / hubble[1:]) / 2. * dz)
        self._z, self._r = z, r

    def get_r_comov(self, z):
        return np.interp(z, self._z, self._r)

    get_dist_m = get_r_comov

    def table(self):
        """(z, r_comov, dist_m) tables (flat: dist_m = r_comov, constants.py:214-215)"""
        return self._z, self._r, self._r


def compute_ang_max(cosmo, r_trans_max, z_min, z_min2=None):
    """reference py/picca/utils.py:419-450"""
    if z_min2 is None:
        z_min2 = z_min
    r_min = cosmo.get_dist_m(z_min)
    r_min2 = cosmo.get_dist_m(z_min2)
    if r_min + r_min2 < r_trans_max:
        return np.pi
    return float(2. * np.arcsin(r_trans_max / (r_min + r_min2)))


def ang2pix_ring(nside, theta, phi):
    """HEALPix RING index of (theta, phi) -- published HEALPix formulae (what healpy.ang2pix
    computes at reference py/picca/io.py:488)."""
    theta = np.atleast_1d(np.asarray(theta, dtype=np.float64))
    phi = np.atleast_1d(np.asarray(phi, dtype=np.float64))
    z = np.cos(theta)
    za = np.abs(z)
    tt = np.mod(phi, 2.0 * np.pi) / (0.5 * np.pi)
    pix = np.empty(z.shape, dtype=np.int64)
    eq = za <= 2.0 / 3.0
    temp1 = nside * (0.5 + tt[eq])
    temp2 = nside * z[eq] * 0.75
    jp = np.floor(temp1 - temp2).astype(np.int64)
    jm = np.floor(temp1 + temp2).astype(np.int64)
    ir = nside + 1 + jp - jm
    kshift = 1 - (ir & 1)
    ip = np.mod((jp + jm - nside + kshift + 1) // 2, 4 * nside)
    pix[eq] = 2 * nside * (nside - 1) + (ir - 1) * 4 * nside + ip
    cap = ~eq
    tp = tt[cap] - np.floor(tt[cap])
    tmp = nside * np.sqrt(3.0 * (1.0 - za[cap]))
    jp = np.floor(tp * tmp).astype(np.int64)
    jm = np.floor((1.0 - tp) * tmp).astype(np.int64)
    ir = jp + jm + 1
    ip = np.mod(np.floor(tt[cap] * ir).astype(np.int64), 4 * ir)
    pix[cap] = np.where(z[cap] > 0, 2 * ir * (ir - 1) + ip,
                        12 * nside * nside - 2 * ir * (ir + 1) + ip)
    return pix


def project(delta, weights, log_lambda, order):
    """Delta.project(), reference py/picca/data.py:622-655."""
    sum_weights = np.sum(weights)
    if not sum_weights > 0.0:
        return delta
    mean_delta = np.average(delta, weights=weights)
    res = 0
    if order == 1 and delta.shape[0] > 1:
        mean_log_lambda = np.average(log_lambda, weights=weights)
        meanless = log_lambda - mean_log_lambda
        res = (np.sum(weights * delta * meanless) / np.sum(weights * meanless**2)) * meanless
    elif order == 1:
        res = delta
    return delta - (mean_delta + res)


def make_forests(n_forest, seed=20260102, nside=32, ra_deg=(0., 120.), dec_deg=(0., 40.2),
                 rest_range=(1045., 1192.), lambda_min=3600., dlambda=0.8, z_ref=2.25, alpha=2.9,
                 order=1, zero_weight_frac=0.02, cosmo=None, id_offset=0, max_pix=None):
    """DR16-like synthetic Lyman-alpha forests.

    Returns (data, num_data, z_min, z_max, cosmo) -- the tuple picca_cf.py builds at :387-405.
    """
    rng = np.random.default_rng(seed)
    cosmo = cosmo or FlatLCDM()
    ra = np.radians(rng.uniform(ra_deg[0], ra_deg[1], n_forest))
    s0, s1 = np.sin(np.radians(dec_deg[0])), np.sin(np.radians(dec_deg[1]))
    dec = np.arcsin(rng.uniform(s0, s1, n_forest))
    z_qso = 2.1 + rng.exponential(0.45, n_forest)
    z_qso = np.where(z_qso > 3.6, 2.1 + (z_qso - 2.1) % 1.5, z_qso)
    healpix = ang2pix_ring(nside, np.pi / 2. - dec, ra)

    k_lo = np.ceil((rest_range[0] * (1. + z_qso) - lambda_min) / dlambda).astype(np.int64)
    k_lo = np.maximum(k_lo, 0)
    k_hi = np.floor((rest_range[1] * (1. + z_qso) - lambda_min) / dlambda).astype(np.int64)
    npix = np.maximum(k_hi - k_lo + 1, 2)
    if max_pix is not None:
        npix = np.minimum(npix, max_pix)

    # the forests are written as views into one array per field, forests back to back in
    # catalogue order (ascending HEALPix, generation order inside a pixel), like the B200 loader
    # does, and registered for catalog.pack (forest.register_soa)
    order_cat = np.argsort(healpix, kind="stable")
    offset = np.zeros(n_forest + 1, dtype=np.int64)
    np.cumsum(npix[order_cat], out=offset[1:])
    start = np.empty(n_forest, dtype=np.int64)
    start[order_cat] = offset[:-1]
    soa = {name: np.empty(int(offset[-1]), dtype=np.float64) for name in PIXEL_FIELDS}
    soa["offset"] = offset
    data = {}
    z_min, z_max = np.inf, 0.
    for f in range(n_forest):
        n = int(npix[f])
        a = int(start[f])
        lam = lambda_min + dlambda * (k_lo[f] + np.arange(n))
        log_lambda = soa["log_lambda"][a:a + n]
        log_lambda[:] = np.log10(lam)
        delta = rng.normal(0., 0.25, n)
        weights = rng.uniform(0.5, 2.0, n)
        weights[rng.random(n) < zero_weight_frac] = 0.
        z = soa["z"][a:a + n]
        z[:] = 10**log_lambda / LYA - 1.
        w_view = soa["weights"][a:a + n]
        w_view[:] = weights * ((1 + z) / (1 + z_ref))**(alpha - 1)  # io.py:503
        d_view = soa["delta"][a:a + n]
        d_view[:] = project(delta, w_view, log_lambda, order)        # io.py:505-506
        tid = id_offset + f + 1
        d = Delta(tid, float(ra[f]), float(dec[f]), float(z_qso[f]), tid, tid, tid, log_lambda,
                  w_view, d_view, order)
        d.z = z
        d.r_comov = soa["r_comov"][a:a + n]
        d.r_comov[:] = cosmo.get_r_comov(z)
        d.dist_m = soa["dist_m"][a:a + n]
        d.dist_m[:] = cosmo.get_dist_m(z)
        z_min = min(z_min, z.min())
        z_max = max(z_max, z.max())
        data.setdefault(int(healpix[f]), []).append(d)
    register_soa(data, soa)
    return data, n_forest, float(z_min), float(z_max), cosmo


def make_quasars(n_qso, seed=20260103, nside=32, ra_deg=(0., 120.), dec_deg=(0., 40.2),
                 z_range=(1.8, 3.6), z_ref=2.25, alpha_obj=1.44, cosmo=None, id_offset=10**7):
    """Synthetic quasar catalogue as ``io.read_objects`` returns it (io.py:595-612):
    (objs, z_min_obj)."""
    rng = np.random.default_rng(seed)
    cosmo = cosmo or FlatLCDM()
    ra = np.radians(rng.uniform(ra_deg[0], ra_deg[1], n_qso))
    s0, s1 = np.sin(np.radians(dec_deg[0])), np.sin(np.radians(dec_deg[1]))
    dec = np.arcsin(rng.uniform(s0, s1, n_qso))
    z = rng.uniform(z_range[0], z_range[1], n_qso)
    healpix = ang2pix_ring(nside, np.pi / 2. - dec, ra)
    objs = {}
    for q in np.argsort(healpix, kind="stable"):
        tid = id_offset + int(q) + 1
        o = QSO(tid, float(ra[q]), float(dec[q]), float(z[q]), tid, tid, tid)
        o.weights = ((1. + o.z_qso) / (1. + z_ref))**(alpha_obj - 1.)
        o.r_comov = float(cosmo.get_r_comov(o.z_qso))
        o.dist_m = float(cosmo.get_dist_m(o.z_qso))
        objs.setdefault(int(healpix[q]), []).append(o)
    return objs, float(z.min())
