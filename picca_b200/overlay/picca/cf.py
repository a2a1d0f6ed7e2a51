"""``picca.cf`` -> ``picca_b200.cf`` (same module object: the scripts assign its globals)."""
import sys

import picca_b200.cf as _impl

sys.modules[__name__] = _impl
