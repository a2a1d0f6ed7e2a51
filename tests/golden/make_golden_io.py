"""Golden vectors of the delta loader: what the LIVE reference's io.read_deltas (imported
unmodified from /root/reference through tests/refharness) returns for (a) the copied reference
fixture tests/golden/fixtures/delta-272.fits.gz and (b) the generated files of cases_io.py, with
the reference's own Cosmo(Om=0.315) tables.

    python -m tests.golden.make_golden_io        # writes tests/golden/golden_io.npz
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from tests.golden import cases_io  # noqa: E402
from tests.refharness import load  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    _, _, io, constants, _ = load.reference_modules()
    io.userprint = lambda *a, **k: None
    cosmo = constants.Cosmo(Om=0.315, Or=0., Ok=0., wl=-1., blinding="none")
    out = {"cosmo_z": cosmo.get_r_comov.x, "cosmo_r_comov": cosmo.get_r_comov.y,
           "cosmo_dist_m": cosmo.get_dist_m.y}

    def run(tag, in_dir, attr, **kw):
        data, num, z_min, z_max = io.read_deltas(in_dir, cosmo=cosmo, nproc=1,
                                                 delta_attributes=attr,
                                                 **dict(cases_io.READ_KW, **kw))
        flat = cases_io.flatten(data)
        for k, v in flat.items():
            out["%s_%s" % (tag, k)] = v
        out["%s_summary" % tag] = np.array([num, z_min, z_max])
        print(tag, num, z_min, z_max, flat["n_pix"].sum())

    fx = os.path.join(HERE, "fixtures")
    run("fixture", os.path.join(fx, "delta-272.fits.gz"),
        os.path.join(fx, "delta_attributes.fits.gz"))
    run("imagefixture", os.path.join(fx, "image-delta-50.fits.gz"),
        os.path.join(fx, "delta_attributes.fits.gz"))
    run("imagefixture_rebin3", os.path.join(fx, "image-delta-50.fits.gz"),
        os.path.join(fx, "delta_attributes.fits.gz"), rebin_factor=3)
    with tempfile.TemporaryDirectory() as tmp:
        for name in cases_io.IMAGE_CASES:
            in_dir, attr = cases_io.write_image_case(tmp, name)
            run(name, in_dir, attr)
        for name in cases_io.CASES:
            in_dir, attr = cases_io.write_case(tmp, name)
            run(name, in_dir, attr)
        in_dir, attr = cases_io.write_case(tmp, "sdss")
        run("sdss_noproject", in_dir, attr, no_project=True)
        run("sdss_max30", in_dir, attr, max_num_spec=30)
        run("sdss_zcut", in_dir, attr, z_min_qso=2.4, z_max_qso=3.0)
        in_dir, attr = cases_io.write_case(tmp, "lin")
        run("lin_rebin2", in_dir, attr, rebin_factor=2)
        in_dir, attr = cases_io.write_image_case(tmp, "image")
        run("image_rebin3", in_dir, attr, rebin_factor=3)
    np.savez_compressed(os.path.join(HERE, "golden_io.npz"), **out)


if __name__ == "__main__":
    main()
