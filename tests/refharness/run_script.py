"""Launch one of the reference's UNMODIFIED scripts in this process, either on the B200 path
(``--impl b200``: ``picca_b200.run``, i.e. the overlay) or on the oracle doubles / the reference's
own Numba modules (CPU).  Test infrastructure: the stand-ins for healpy / fitsio / astropy that
this image lacks are installed first (tests/refharness/shims.py); a site with the real packages
runs ``python -m picca_b200.run picca_cf.py ...`` directly.

    python -m tests.refharness.run_script --impl b200|oracle|reference picca_cf.py <script args>
"""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def main(argv):
    assert argv[0] == "--impl"
    impl, script, args = argv[1], argv[2], argv[3:]
    from tests.refharness import shims
    assert shims.install(), "no reference tree (neither /root/reference nor baseline/_ref)"
    name = script[:-3] if script.endswith(".py") else script
    if impl == "b200":
        from picca_b200 import run
        run.main([script] + args)
        return
    mod = importlib.import_module("picca.bin." + name)
    if impl == "oracle":
        for double in ("cf", "xcf", "co"):
            if hasattr(mod, double):
                setattr(mod, double, importlib.import_module("oracle." + double))
    mod.main(args)


if __name__ == "__main__":
    main(sys.argv[1:])
