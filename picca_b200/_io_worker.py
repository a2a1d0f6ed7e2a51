"""Child process of the isolated delta loader (``picca_b200.io.read_deltas(isolate=True)``):

    python -m picca_b200._io_worker <exchange dir>

Reads ``args.pkl`` from the exchange directory, runs the device loader, writes every SoA array
as ``<name>.npy`` and the scalars (or the exception) as ``meta.pkl``.  The parent never touches
CUDA, so the reference's scripts can fork their worker pools afterwards (picca_cf.py:455)."""
import os
import pickle
import sys

import numpy as np


def main(tmp):
    from picca_b200 import io as pio
    with open(os.path.join(tmp, "args.pkl"), "rb") as f:
        args, kwds = pickle.load(f)
    try:
        soa = pio.read_deltas_soa(*args, **kwds)
        arrays = [k for k in soa if isinstance(soa[k], np.ndarray)]
        for name in arrays:
            np.save(os.path.join(tmp, name + ".npy"), soa[name])
        meta = {"arrays": arrays,
                "scalars": {k: v for k, v in soa.items() if not isinstance(v, np.ndarray)}}
    except Exception as err:  # re-raised by the parent with the same type
        meta = {"error": (type(err).__name__, str(err))}
    with open(os.path.join(tmp, "meta.pkl.tmp"), "wb") as f:
        pickle.dump(meta, f)
    os.replace(os.path.join(tmp, "meta.pkl.tmp"), os.path.join(tmp, "meta.pkl"))


if __name__ == "__main__":
    main(sys.argv[1])
