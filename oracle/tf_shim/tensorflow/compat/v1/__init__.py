"""tf.compat.v1: only the legacy image resize the reference calls
(ddsp/core.py:617-620)."""
import types as _types
import tensorflow as _tf

image = _types.SimpleNamespace(resize=_tf._v1_resize, ResizeMethod=_tf._ResizeMethod,
                               resize_images=_tf._v1_resize)
