"""Oracle restatement of the delta loader: ``io.read_deltas`` / ``io.read_delta_file`` /
``Delta.from_fitsio`` / ``Delta.project`` (reference py/picca/io.py:338-381, :383-512;
py/picca/data.py:375-474, :519-620, :622-655), BinTable and ImageHDU flavours, NumPy + scipy on
the CPU.
TEST INFRASTRUCTURE ONLY -- the referee for picca_b200.io, never the product.

Pinned bit for bit against the live reference on its bundled delta files and on the generated
cases (tests/test_oracle_vs_reference.py) and against tests/golden/golden_io.npz.  FITS access
goes through the harness reader (tests/refharness/minifits.py), the stand-in for fitsio.
"""
import glob
import os

import numpy as np
from scipy import interpolate

from picca_b200.forest import Delta
from picca_b200.synth import ang2pix_ring
from tests.refharness import minifits


def find_order(delta_attributes):
    """io.py:53-56 (FIT_METADATA) and :64-68 (deprecated STACK_DELTAS location)"""
    with minifits.FITS(delta_attributes) as hdul:
        for ext in ("FIT_METADATA", "STACK_DELTAS"):
            if ext in hdul and "FITORDER" in hdul[ext].read_header():
                return hdul[ext].read_header()["FITORDER"]
    return None


def project(d):
    """data.py:622-655"""
    sum_weights = np.sum(d.weights)
    if not sum_weights > 0.0:
        return
    mean_delta = np.average(d.delta, weights=d.weights)
    res = 0
    if d.order == 1 and d.delta.shape[0] > 1:
        mean_log_lambda = np.average(d.log_lambda, weights=d.weights)
        meanless = d.log_lambda - mean_log_lambda
        res = (np.sum(d.weights * d.delta * meanless) /
               np.sum(d.weights * meanless**2)) * meanless
    elif d.order == 1:
        res = d.delta
    d.delta -= mean_delta + res


def from_image(hdul, z_min_qso, z_max_qso, order):
    """Delta.from_image, data.py:543-620 (non-Pk1D)"""
    meta = hdul["METADATA"]
    header = meta.read_header()
    blinding = header["BLINDING"] if "BLINDING" in header else "none"
    delta = hdul["DELTA" if blinding == "none" else "DELTA_BLIND"].read().astype(float)
    if "LOGLAM" in hdul:
        log_lambda = hdul["LOGLAM"][:].astype(float)
    else:
        log_lambda = np.log10(hdul["LAMBDA"][:].astype(float))
    weights = hdul["WEIGHT"].read().astype(float)
    keep = weights > 0                                         # :572
    if "THING_ID" in meta.get_colnames():
        ids = [meta[c][:] for c in ("THING_ID", "PLATE", "MJD", "FIBERID")]
    else:
        ids = [meta["LOS_ID"][:]] * 4
    ra, dec, z_qso = meta["RA"][:], meta["DEC"][:], meta["Z"][:]
    out = []
    for f in range(len(z_qso)):
        if z_qso[f] >= z_min_qso and z_qso[f] <= z_max_qso:   # :602, inclusive
            w = keep[f]
            out.append(Delta(ids[0][f], ra[f], dec[f], z_qso[f], ids[1][f], ids[2][f], ids[3][f],
                             log_lambda[w], weights[f][w], delta[f][w], order))
    return out


def rebin(d, factor, dwave):
    """Delta.rebin, data.py:666-686"""
    wave = 10**np.array(d.log_lambda)
    start = wave.min() - dwave / 2
    num_bins = np.ceil(((wave[-1] - wave[0]) / dwave + 1) / factor)
    edges = np.arange(num_bins) * dwave * factor + start
    new_indx = np.searchsorted(edges, wave)
    binned_delta = np.bincount(new_indx, weights=d.delta * d.weights,
                               minlength=edges.size + 1)[1:-1]
    binned_weight = np.bincount(new_indx, weights=d.weights, minlength=edges.size + 1)[1:-1]
    mask = binned_weight != 0
    binned_delta[mask] /= binned_weight[mask]
    new_wave = (edges[1:] + edges[:-1]) / 2
    d.log_lambda = np.log10(new_wave[mask])
    d.delta = binned_delta[mask]
    d.weights = binned_weight[mask]


def read_delta_file(filename, z_min_qso, z_max_qso, order, rebin_factor=None):
    """io.py:354-380 + data.py:392-474 (non-Pk1D branch)"""
    out = []
    with minifits.FITS(filename) as hdul:
        if rebin_factor is not None:                           # io.py:362-373
            head = hdul["LAMBDA" if "LAMBDA" in hdul else 1].read_header()
            if head["WAVE_SOLUTION"] != "lin":
                raise ValueError("Delta rebinning only implemented for linear lambda bins")
            dwave = head["DELTA_LAMBDA"]
        if "LAMBDA" in hdul:                                   # io.py:356
            out = from_image(hdul, z_min_qso, z_max_qso, order)
            for d in out if rebin_factor is not None else ():
                rebin(d, rebin_factor, dwave)
            return out
        for hdu in hdul[1:]:
            header = hdu.read_header()
            if not z_min_qso < header["Z"] < z_max_qso:
                continue
            blinding = header["BLINDING"] if "BLINDING" in header else "none"
            delta = hdu["DELTA" if blinding == "none" else "DELTA_BLIND"][:].astype(float)
            if "LOGLAM" in hdu.get_colnames():
                log_lambda = hdu["LOGLAM"][:].astype(float)
            else:
                log_lambda = np.log10(hdu["LAMBDA"][:].astype(float))
            weights = hdu["WEIGHT"][:].astype(float)
            if "THING_ID" in header:
                ids = (header["THING_ID"], header["PLATE"], header["MJD"], header["FIBERID"])
            else:
                ids = (header["LOS_ID"],) * 4
            out.append(Delta(ids[0], header["RA"], header["DEC"], header["Z"], ids[1], ids[2],
                             ids[3], log_lambda, weights, delta, order))
    for d in out if rebin_factor is not None else ():
        rebin(d, rebin_factor, dwave)
    return out


def read_deltas(in_dir, nside, lambda_abs, alpha, z_ref, tables, max_num_spec=None,
                no_project=False, z_min_qso=0, z_max_qso=10, delta_attributes=None,
                rebin_factor=None):
    """io.py:448-512.  ``tables`` = (z, r_comov, dist_m) of the cosmology; the distances are
    evaluated with scipy's interp1d like constants.py:211-229."""
    if in_dir.endswith(".fits.gz") or in_dir.endswith(".fits"):
        files = sorted(glob.glob(in_dir))
    else:
        files = sorted(glob.glob(in_dir + "/*.fits") + glob.glob(in_dir + "/*.fits.gz"))
    order = find_order(delta_attributes)
    deltas = []
    for f in files:
        deltas += read_delta_file(f, z_min_qso, z_max_qso, order, rebin_factor)
        if max_num_spec is not None and len(deltas) > max_num_spec:
            break
    if max_num_spec is not None:
        deltas = deltas[:max_num_spec]
    get_r_comov = interpolate.interp1d(tables[0], tables[1])
    get_dist_m = interpolate.interp1d(tables[0], tables[2])
    healpixs = ang2pix_ring(nside, np.pi / 2. - np.array([d.dec for d in deltas]),
                            np.array([d.ra for d in deltas]))
    data, z_min, z_max = {}, None, 0.
    for d, hp in zip(deltas, healpixs):
        z = 10**d.log_lambda / lambda_abs - 1.
        z_min = z.min() if z_min is None else min(z_min, z.min())
        z_max = max(z_max, z.max())
        d.z = z
        d.r_comov = get_r_comov(z)
        d.dist_m = get_dist_m(z)
        d.weights *= ((1 + z) / (1 + z_ref))**(alpha - 1)
        if not no_project:
            project(d)
        data.setdefault(int(hp), []).append(d)
    return data, len(deltas), z_min, z_max
