"""tensorflow_probability is imported at module level by ddsp/losses.py and
spectral_ops.py (class bases, aliases); nothing on the decoder path calls it."""


class _Namespace:
  def __getattr__(self, item):
    if item.startswith('__'):
      raise AttributeError(item)
    return type(item, (), {'__init__': lambda self, *a, **k: (_ for _ in ()).throw(
        NotImplementedError('tfp.%s is not available (tf_shim stub)' % item))})


distributions = _Namespace()
