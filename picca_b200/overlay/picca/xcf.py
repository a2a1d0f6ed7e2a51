"""``picca.xcf`` -> ``picca_b200.xcf`` (same module object: the scripts assign its globals)."""
import sys

import picca_b200.xcf as _impl

sys.modules[__name__] = _impl
