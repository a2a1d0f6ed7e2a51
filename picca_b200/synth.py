"""Deterministic synthetic forests / quasars of the shapes BASELINE.json names.

The generator produces exactly what ``picca.io.read_deltas`` / ``read_objects`` hand to the hot
path (reference py/picca/io.py:485-512, :595-612): ``dict[healpix] -> list[Delta]`` with z, r_comov,
dist_m filled, weights scaled by ((1+z)/(1+z_ref))**(alpha-1) and deltas projected.  It does not
read or need the reference.
"""
from concurrent.futures import ThreadPoolExecutor

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


# ---------------------------------------------------------------------------------------------
# Banded generation (DESI-scale surveys on several GPUs): the positions and shapes of ALL forests
# are drawn first -- a light index, vectorised -- and the pixel data only for the band of HEALPix
# rows a rank holds (dist.BandShard: its band + halo).  The pixel content of a HEALPix row is
# drawn from a generator seeded with (seed, healpix), so two ranks that both hold a row (one as
# halo) hold the same forests.  Same shapes, weights and projection as make_forests; a different
# random stream (make_forests draws forest by forest).
# ---------------------------------------------------------------------------------------------
class ForestIndex:
    """Every forest of a synthetic survey in catalogue order (ascending HEALPix, generation order
    inside a pixel): position, quasar redshift, first wavelength bin and pixel count."""

    def __init__(self, **kw):
        self.__dict__.update(kw)


def make_forest_index(n_forest, seed=20260102, nside=32, ra_deg=(0., 120.), dec_deg=(0., 40.2),
                      rest_range=(1045., 1192.), lambda_min=3600., dlambda=0.8, z_ref=2.25,
                      alpha=2.9, order=1, zero_weight_frac=0.02, cosmo=None, max_pix=None):
    rng = np.random.default_rng(seed)
    cosmo = cosmo or FlatLCDM()
    ra = np.radians(rng.uniform(ra_deg[0], ra_deg[1], n_forest))
    s0, s1 = np.sin(np.radians(dec_deg[0])), np.sin(np.radians(dec_deg[1]))
    dec = np.arcsin(rng.uniform(s0, s1, n_forest))
    z_qso = 2.1 + rng.exponential(0.45, n_forest)
    z_qso = np.where(z_qso > 3.6, 2.1 + (z_qso - 2.1) % 1.5, z_qso)
    healpix = ang2pix_ring(nside, np.pi / 2. - dec, ra)
    k_lo = np.maximum(np.ceil((rest_range[0] * (1. + z_qso) - lambda_min) / dlambda).astype(np.int64), 0)
    k_hi = np.floor((rest_range[1] * (1. + z_qso) - lambda_min) / dlambda).astype(np.int64)
    npix = np.maximum(k_hi - k_lo + 1, 2)
    if max_pix is not None:
        npix = np.minimum(npix, max_pix)
    cat = np.argsort(healpix, kind="stable")
    healpixs, counts = np.unique(healpix, return_counts=True)
    first = np.zeros(healpixs.size + 1, dtype=np.int64)
    np.cumsum(counts, out=first[1:])
    ra, dec, z_qso, k_lo, npix = ra[cat], dec[cat], z_qso[cat], k_lo[cat], npix[cat]
    lam_lo = lambda_min + dlambda * k_lo
    lam_hi = lambda_min + dlambda * (k_lo + npix - 1)
    return ForestIndex(
        n_forest=int(n_forest), seed=int(seed), nside=nside, healpixs=[int(h) for h in healpixs],
        counts=counts.astype(np.int64), first=first, ra=ra, dec=dec, z_qso=z_qso, k_lo=k_lo,
        npix=npix.astype(np.int64), cosmo=cosmo, lambda_min=lambda_min, dlambda=dlambda,
        z_ref=z_ref, alpha=alpha, order=order, zero_weight_frac=zero_weight_frac,
        z_min=float((10**np.log10(lam_lo) / LYA - 1.).min()),
        z_max=float((10**np.log10(lam_hi) / LYA - 1.).max()),
        xyz=np.stack([np.cos(ra) * np.cos(dec), np.sin(ra) * np.cos(dec), np.sin(dec)], axis=1))


def make_forest_band(index, r0, r1, rows_per_pass=32, threads=2):
    """``data`` dict (registered SoA, like make_forests) of the HEALPix rows
    ``index.healpixs[r0:r1]``; thingid = 1 + position in the whole catalogue."""
    ix = index
    f0, f1 = int(ix.first[r0]), int(ix.first[r1])
    nf = f1 - f0
    npix = ix.npix[f0:f1]
    offset = np.zeros(nf + 1, dtype=np.int64)
    np.cumsum(npix, out=offset[1:])
    total = int(offset[-1])
    soa = {name: np.empty(total, dtype=np.float64) for name in PIXEL_FIELDS}
    soa["offset"] = offset
    kmax = int((ix.k_lo + ix.npix).max())
    tab = {"log_lambda": np.log10(ix.lambda_min + ix.dlambda * np.arange(kmax + 1))}
    tab["z"] = 10**tab["log_lambda"] / LYA - 1.
    tab["wscale"] = ((1 + tab["z"]) / (1 + ix.z_ref))**(ix.alpha - 1)
    tab["r_comov"] = ix.cosmo.get_r_comov(tab["z"])
    def one_pass(ra_):
        rb_ = min(r1, ra_ + rows_per_pass)
        a, b = int(ix.first[ra_]) - f0, int(ix.first[rb_]) - f0       # forests of the pass
        pa, pb = int(offset[a]), int(offset[b])                       # their pixels
        delta = soa["delta"][pa:pb]
        weights = soa["weights"][pa:pb]
        for r in range(ra_, rb_):   # pixel content: one generator per HEALPix row
            qa, qb = int(offset[int(ix.first[r]) - f0]) - pa, int(offset[int(ix.first[r + 1]) - f0]) - pa
            rng = np.random.default_rng([ix.seed, ix.healpixs[r]])
            delta[qa:qb] = rng.normal(0., 0.25, qb - qa)
            w = rng.uniform(0.5, 2.0, qb - qa)
            w[rng.random(qb - qa) < ix.zero_weight_frac] = 0.
            weights[qa:qb] = w
        n = npix[a:b]
        fid = np.repeat(np.arange(b - a, dtype=np.int32), n)
        start = offset[a:b] - pa
        # every per-pixel quantity below is a function of the wavelength bin only: tables over
        # the bins, formed with the expressions of make_forests, and one gather per field
        kbin = (ix.k_lo[f0 + a:f0 + b] - start)[fid] + np.arange(pb - pa)
        log_lambda = soa["log_lambda"][pa:pb]
        np.take(tab["log_lambda"], kbin, out=log_lambda)
        z = soa["z"][pa:pb]
        np.take(tab["z"], kbin, out=z)
        weights *= tab["wscale"][kbin]                                    # io.py:503
        np.take(tab["r_comov"], kbin, out=soa["r_comov"][pa:pb])
        soa["dist_m"][pa:pb] = soa["r_comov"][pa:pb]
        # Delta.project (data.py:622-655), all forests of the pass at once
        sw = np.add.reduceat(weights, start)
        sw = np.where(sw > 0, sw, 1.)
        mean_delta = np.add.reduceat(weights * delta, start) / sw
        if ix.order == 1:
            meanless = log_lambda - (np.add.reduceat(weights * log_lambda, start) / sw)[fid]
            den = np.add.reduceat(weights * meanless**2, start)
            slope = np.add.reduceat(weights * delta * meanless, start) / np.where(den > 0, den, 1.)
            delta -= mean_delta[fid] + slope[fid] * meanless
        else:
            delta -= mean_delta[fid]

    # the passes write disjoint slices; NumPy releases the GIL inside its loops
    with ThreadPoolExecutor(max(1, int(threads))) as pool:
        list(pool.map(one_pass, range(r0, r1, rows_per_pass)))
    data = {}
    k = 0
    for r in range(r0, r1):
        row = []
        for _ in range(int(ix.counts[r])):
            a, b = int(offset[k]), int(offset[k + 1])
            tid = f0 + k + 1
            d = Delta(tid, float(ix.ra[f0 + k]), float(ix.dec[f0 + k]), float(ix.z_qso[f0 + k]), tid,
                      tid, tid, soa["log_lambda"][a:b], soa["weights"][a:b], soa["delta"][a:b],
                      ix.order)
            d.z = soa["z"][a:b]
            d.r_comov = soa["r_comov"][a:b]
            d.dist_m = soa["dist_m"][a:b]
            row.append(d)
            k += 1
        data[ix.healpixs[r]] = row
    register_soa(data, soa)
    return data
