"""`import tensorflow.compat.v2 as tf`: the same namespace as `tensorflow`."""
import sys as _sys
import tensorflow as _tf

_me = _sys.modules[__name__]
for _k, _v in list(vars(_tf).items()):
  if not _k.startswith('__'):
    setattr(_me, _k, _v)


def enable_v2_behavior():
  pass
