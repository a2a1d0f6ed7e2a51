"""Deterministic delta files for the loader parity tests: written with the test harness's FITS
writer (tests/refharness/minifits.py) in the layout of SURVEY.md Appendix C, so they can be
regenerated on the GPU box; the live reference's read_deltas output for them is in
golden_io.npz."""
import os

import numpy as np

from tests.refharness import minifits

LYA = 1215.67
NSIDE = 16
READ_KW = dict(nside=NSIDE, lambda_abs=LYA, alpha=2.9, z_ref=2.25)

CASES = {
    # directory name: flavour of the headers / wavelength column, forests per file
    "sdss": dict(seed=21, files=[40, 7, 25], ids="THING_ID", wave="LOGLAM", order=1),
    "desi": dict(seed=22, files=[33, 12], ids="LOS_ID", wave="LAMBDA", order=0),
    "blind": dict(seed=23, files=[9], ids="LOS_ID", wave="LOGLAM", order=1, blinding="desi_m2"),
    # for rebin_factor: no single-pixel and no all-zero-weight forests (rebinning empties them and
    # the reference then fails on the empty arrays, io.py:493-497)
    "lin": dict(seed=24, files=[20, 9], ids="THING_ID", wave="LOGLAM", order=1, min_pix=12,
                keep_weights=True),
}


def write_case(root, name):
    """Write the delta files + delta_attributes of case ``name`` under ``root``; returns
    (in_dir, delta_attributes path)."""
    cfg = CASES[name]
    rng = np.random.default_rng(cfg["seed"])
    in_dir = os.path.join(root, name, "Delta")
    os.makedirs(in_dir, exist_ok=True)
    los = 1000 * cfg["seed"]
    for k, n_forest in enumerate(cfg["files"]):
        out = minifits.FITS(os.path.join(in_dir, "delta-%d.fits.gz" % (100 + k)), "rw",
                            clobber=True)
        for f in range(n_forest):
            los += 1
            z_qso = float(rng.uniform(2.0, 3.5))
            if f % 13 == 5:
                z_qso = 10.5  # outside the default quasar redshift cut (io.py:359-360)
            lo = cfg.get("min_pix", 1)
            n = int(rng.integers(lo, 260)) if f % 7 else lo  # single-pixel forests too
            lam0 = 1040. * (1. + min(z_qso, 3.5)) + rng.uniform(0, 0.8)
            lam = lam0 + 0.8 * np.arange(n)
            delta = rng.normal(0., 0.3, n)
            weight = rng.uniform(0.2, 3., n)
            weight[rng.random(n) < 0.05] = 0.
            if f % 11 == 3 and not cfg.get("keep_weights"):
                weight[:] = 0.  # project() returns early (data.py:636-640)
            cont = rng.uniform(0.5, 2., n)
            wave = np.log10(lam) if cfg["wave"] == "LOGLAM" else lam
            head = [{"name": "RA", "value": float(rng.uniform(0.1, 0.2))},
                    {"name": "DEC", "value": float(rng.uniform(-0.02, 0.08))},
                    {"name": "Z", "value": z_qso},
                    {"name": "PMF", "value": "1-2-3"}]
            if cfg["ids"] == "THING_ID":
                head += [{"name": "THING_ID", "value": los}, {"name": "PLATE", "value": 3000 + f},
                         {"name": "MJD", "value": 55000 + f},
                         {"name": "FIBERID", "value": 1 + (37 * f) % 1000}]
            else:
                head += [{"name": "LOS_ID", "value": 39627000000000000 + los}]
            head.append({"name": "ORDER", "value": cfg["order"]})
            head += [{"name": k, "value": v} for k, v in WAVE_CARDS]
            names = [cfg["wave"], "DELTA", "WEIGHT", "CONT"]
            if "blinding" in cfg:
                head.append({"name": "BLINDING", "value": cfg["blinding"]})
                names[1] = "DELTA_BLIND"
            out.write([wave, delta, weight, cont], names=names, header=head, extname=str(los))
        out.close()
        import gzip
        path = os.path.join(in_dir, "delta-%d.fits.gz" % (100 + k))
        with gzip.open(path, "rb") as fh:
            raw = fh.read()
        with gzip.open(path, "wb") as fh:
            fh.write(_hierarch(raw))
    attr = os.path.join(root, name, "delta_attributes.fits.gz")
    out = minifits.FITS(attr, "rw", clobber=True)
    out.write([np.arange(3.)], names=["LOGLAM"], extname="STACK_DELTAS")
    out.write([np.arange(2.)], names=["X"], header=[{"name": "FITORDER", "value": cfg["order"]}],
              extname="FIT_METADATA")
    out.close()
    return in_dir, attr


def _hierarch(raw):
    """swap the placeholder cards HWAVESOL / HDLAMBDA for the HIERARCH cards fitsio writes for the
    long keywords WAVE_SOLUTION / DELTA_LAMBDA (delta files of picca's delta extraction)"""
    raw = bytes(raw)
    for short, long in ((b"HWAVESOL", b"WAVE_SOLUTION"), (b"HDLAMBDA", b"DELTA_LAMBDA")):
        pos = raw.find(short + b"= ")
        while pos >= 0:
            value = raw[pos + 10:pos + 80].strip()
            card = (b"HIERARCH " + long + b" = " + value).ljust(80)
            raw = raw[:pos] + card + raw[pos + 80:]
            pos = raw.find(short + b"= ")
    return raw


WAVE_CARDS = [("HWAVESOL", "lin"), ("HDLAMBDA", 0.8)]


def _image_hdu(name, arr, extra=()):
    arr = np.asarray(arr, dtype=np.float64)
    cards = [("XTENSION", "IMAGE"), ("BITPIX", -64), ("NAXIS", arr.ndim)]
    cards += [("NAXIS%d" % (k + 1), n) for k, n in enumerate(arr.shape[::-1])]
    cards += [("PCOUNT", 0), ("GCOUNT", 1), ("EXTNAME", name)] + list(extra)
    raw = arr.astype(">f8").tobytes()
    return _hierarch(minifits._cards_to_bytes(cards)) + raw + b"\0" * ((-len(raw)) % minifits.BLOCK)


IMAGE_CASES = {
    "image": dict(seed=31, files=[30, 11], wave="LAMBDA", order=1),
    # (an ImageHDU file with a LOGLAM grid is not reachable through io.read_delta_file: the
    # flavour test is `'LAMBDA' in hdul`, io.py:356)
    "imageblind": dict(seed=32, files=[17], wave="LAMBDA", order=0, blinding="desi_y3"),
}


def write_image_case(root, name):
    """ImageHDU flavour: LAMBDA/LOGLAM grid, METADATA table, DELTA (or
    DELTA_BLIND), WEIGHT, CONT images.  Returns (in_dir, delta_attributes path)."""
    import gzip
    cfg = IMAGE_CASES[name]
    rng = np.random.default_rng(cfg["seed"])
    in_dir = os.path.join(root, name, "Delta")
    os.makedirs(in_dir, exist_ok=True)
    n_lambda = 300
    lam = 3600. + 0.8 * np.arange(n_lambda)
    los = 2000 * cfg["seed"]
    for k, n_forest in enumerate(cfg["files"]):
        z_qso = rng.uniform(2.0, 3.4, n_forest)
        z_qso[::9] = 11.  # outside the default quasar redshift cut (data.py:602, inclusive)
        delta = rng.normal(0., 0.3, (n_forest, n_lambda))
        weight = rng.uniform(0.2, 3., (n_forest, n_lambda))
        for f in range(n_forest):  # a forest covers part of the grid; zeros / negatives elsewhere
            a = int(rng.integers(0, n_lambda - 40))
            b = int(rng.integers(a + 12, n_lambda))
            weight[f, :a] = 0.
            weight[f, b:] = -1.
            weight[f, rng.random(n_lambda) < 0.04] = 0.
        ids = 39627000000000000 + los + np.arange(n_forest)
        los += n_forest
        out = bytearray(minifits._cards_to_bytes([("SIMPLE", True), ("BITPIX", 8), ("NAXIS", 0),
                                                  ("EXTEND", True)]))
        out += _image_hdu(cfg["wave"], lam if cfg["wave"] == "LAMBDA" else np.log10(lam),
                          extra=WAVE_CARDS)
        head = [{"name": "BLINDING", "value": cfg.get("blinding", "none")}]
        out += minifits._table_bytes(
            [ids, rng.uniform(0.1, 0.2, n_forest), rng.uniform(-0.02, 0.08, n_forest), z_qso,
             rng.uniform(1., 5., n_forest), ids, np.full(n_forest, 20210101), np.arange(n_forest),
             np.full(n_forest, 80000 + k)],
            ["LOS_ID", "RA", "DEC", "Z", "MEANSNR", "TARGETID", "NIGHT", "PETAL", "TILE"], None,
            head, "METADATA")
        out += _image_hdu("DELTA_BLIND" if "blinding" in cfg else "DELTA", delta)
        out += _image_hdu("WEIGHT", weight)
        out += _image_hdu("CONT", np.ones((n_forest, n_lambda)))
        with gzip.open(os.path.join(in_dir, "delta-%d.fits.gz" % (200 + k)), "wb") as fh:
            fh.write(bytes(out))
    attr = os.path.join(root, name, "delta_attributes.fits.gz")
    out = minifits.FITS(attr, "rw", clobber=True)
    out.write([np.arange(2.)], names=["X"], header=[{"name": "FITORDER", "value": cfg["order"]}],
              extname="FIT_METADATA")
    out.close()
    return in_dir, attr


def flatten(data):
    """Concatenate a read_deltas dict in a canonical order (healpix, then list order)."""
    out = {k: [] for k in ("healpix", "los_id", "ra", "dec", "z_qso", "plate", "mjd", "fiberid",
                           "n_pix", "order")}
    arrays = {k: [] for k in ("log_lambda", "weights", "delta", "z", "r_comov", "dist_m")}
    for hp in sorted(data):
        for d in data[hp]:
            out["healpix"].append(hp)
            out["los_id"].append(int(d.los_id))
            out["ra"].append(d.ra), out["dec"].append(d.dec), out["z_qso"].append(d.z_qso)
            out["plate"].append(int(d.plate)), out["mjd"].append(int(d.mjd))
            out["fiberid"].append(int(d.fiberid))
            out["n_pix"].append(len(d.weights))
            out["order"].append(-1 if d.order is None else int(d.order))
            for k in arrays:
                arrays[k].append(np.asarray(getattr(d, k), dtype=np.float64))
    res = {k: np.array(v) for k, v in out.items()}
    res.update({k: np.concatenate(v) if v else np.zeros(0) for k, v in arrays.items()})
    return res
