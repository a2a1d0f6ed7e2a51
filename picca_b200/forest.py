"""Line-of-sight records with the attribute names the reference's data model uses.

The product consumes ``dict[healpix] -> list[Delta]`` / ``list[QSO]`` exactly as
``picca.io.read_deltas`` / ``read_objects`` produce them (reference ``py/picca/data.py:14-162`` for
``QSO``, ``:238-373`` for ``Delta``).  When the reference's own classes are available they are used
as-is (duck typing); these light-weight stand-ins carry the same attributes so that the synthetic
generator, the tests and the benchmark can build inputs where the reference is not installed.
"""
import numpy as np

# reference py/picca/constants.py:16 -- 2 arcsec
SMALL_ANGLE_CUT_OFF = 2. / 3600. * np.pi / 180.


class QSO:
    """Attributes as reference ``data.QSO`` (``py/picca/data.py:60-104``)."""

    def __init__(self, los_id, ra, dec, z_qso, plate, mjd, fiberid):
        self.ra = ra
        self.dec = dec
        self.plate = plate
        self.mjd = mjd
        self.fiberid = fiberid
        self.x_cart = np.cos(ra) * np.cos(dec)
        self.y_cart = np.sin(ra) * np.cos(dec)
        self.z_cart = np.sin(dec)
        self.cos_dec = np.cos(dec)
        self.z_qso = z_qso
        self.los_id = los_id
        self.thingid = los_id
        self.weights = None
        self.r_comov = None
        self.dist_m = None
        self.log_lambda = None
        self.neighbours = None


class Delta(QSO):
    """Attributes as reference ``data.Delta`` (``py/picca/data.py:296-373``), hot-path subset."""

    def __init__(self, los_id, ra, dec, z_qso, plate, mjd, fiberid, log_lambda, weights, delta,
                 order):
        QSO.__init__(self, los_id, ra, dec, z_qso, plate, mjd, fiberid)
        self.log_lambda = log_lambda
        self.weights = weights
        self.delta = delta
        self.order = order
        self.z = None
        self.r_comov = None
        self.dist_m = None
        self.neighbours = None
