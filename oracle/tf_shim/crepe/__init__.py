"""crepe is imported by ddsp/losses.py and ddsp/spectral_ops.py at module level
but never touched by the decoder / SpectralLoss path."""


class _Missing:
  def __getattr__(self, item):
    raise NotImplementedError('crepe is not available (tf_shim stub)')


core = _Missing()


def predict(*args, **kwargs):
  raise NotImplementedError('crepe is not available (tf_shim stub)')
