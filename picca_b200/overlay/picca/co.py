"""``picca.co`` -> ``picca_b200.co`` (same module object: the script assigns its globals)."""
import sys

import picca_b200.co as _impl

sys.modules[__name__] = _impl
