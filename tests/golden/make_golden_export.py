"""Golden vectors for the covariance step: outputs of the LIVE reference's utils.compute_cov /
utils.smooth_cov (imported unmodified from /root/reference through tests/refharness) on seeded
inputs, plus the reference's own fixture pair cf.fits.gz -> exported_cf.fits.gz (the DA/WE/RP/RT
inputs and the CO column its test compares at rtol 1e-5).

    python -m tests.golden.make_golden_export      # writes tests/golden/golden_export.npz
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from tests.golden import cases_export  # noqa: E402
from tests.refharness import load, minifits  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    _, _, _, _, utils = load.reference_modules()
    utils.userprint = lambda *a, **k: None
    out = {}
    for name, cfg in cases_export.CASES.items():
        xi, we, rp, rt = cases_export.inputs(cfg)
        cov = utils.compute_cov(xi, we)
        out["%s_cov" % name] = cov
        out["%s_smooth" % name] = utils.smooth_cov(
            xi, we, rp, rt, delta_r_trans=cfg["delta_r_trans"], delta_r_par=cfg["delta_r_par"],
            covariance=cov.copy(), per_r_par=cfg.get("per_r_par", False))
        with np.errstate(all="ignore"):
            out["%s_boot" % name] = utils.compute_cov_boot(xi, we, nboots=cases_export.NBOOTS,
                                                           seed=cases_export.BOOT_SEED)
        print(name, cov.shape, float(np.trace(cov)))
    # the reference's own fixtures (picca_export.py --data cf.fits.gz, test_3_cor.py:443-456)
    cor = load.DATA + "/test_cor/"
    h = minifits.FITS(cor + "cf.fits.gz")
    head = h[1].read_header()
    out["fixture_rp"] = np.array(h[1]["RP"][:])
    out["fixture_rt"] = np.array(h[1]["RT"][:])
    out["fixture_da"] = np.array(h[2]["DA"][:])
    out["fixture_we"] = np.array(h[2]["WE"][:])
    out["fixture_bins"] = np.array([head["NP"], head["NT"], head["RPMIN"], head["RPMAX"],
                                    head["RTMAX"]], dtype=np.float64)
    h.close()
    h = minifits.FITS(cor + "exported_cf.fits.gz")
    out["fixture_co"] = np.array(h[1]["CO"][:])
    out["fixture_exported_da"] = np.array(h[1]["DA"][:])
    h.close()
    np.savez_compressed(os.path.join(HERE, "golden_export.npz"), **out)


if __name__ == "__main__":
    main()
