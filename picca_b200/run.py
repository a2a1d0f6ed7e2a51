"""Run one of the reference's UNMODIFIED correlation scripts on the B200 path:

    python -m picca_b200.run picca_cf.py   --in-dir ... --out cf.fits.gz --nproc 8 ...
    python -m picca_b200.run picca_dmat.py --in-dir ... --out dmat.fits.gz --rej 0.99 ...
    python -m picca_b200.run picca_xcf.py / picca_xdmat.py ...

The script, its argument parsing, its I/O and its fork pool are the reference's own
(py/picca/bin/picca_cf.py etc.); only ``picca.cf`` / ``picca.xcf`` resolve to picca_b200.  Each
forked pool worker binds one GPU (worker index modulo the number of visible devices).
"""
import importlib
import sys

from . import overlay


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        raise SystemExit(__doc__)
    script = argv.pop(0)
    name = script[:-3] if script.endswith(".py") else script
    overlay.activate()
    mod = importlib.import_module("picca.bin." + name)
    mod.main(argv)


if __name__ == "__main__":
    main()
