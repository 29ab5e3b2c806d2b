"""tf.compat namespace of the NumPy shim (see ../__init__.py)."""
import sys as _sys
from tensorflow.compat import v1, v2  # noqa: F401

v2.compat = _sys.modules[__name__]
