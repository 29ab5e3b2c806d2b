"""Runs the UNMODIFIED reference (`/root/reference/ddsp`) on the NumPy
TensorFlow shim (`oracle/tf_shim`).  TEST INFRASTRUCTURE ONLY.

    from oracle import ref_on_shim
    ddsp = ref_on_shim.load()            # the reference package itself
    audio = ddsp.synths.Harmonic(n_samples=16000)(amps, hd, f0)   # tf_shim.Tensor

The reference is pure Python over TensorFlow; TensorFlow cannot be installed in
this image, so `tensorflow`, `gin`, `crepe`, `librosa` and
`tensorflow_probability` resolve to the stand-ins under oracle/tf_shim (only
`tensorflow` and `gin` matter on the decoder path).  No reference source is
copied: the package is imported from where it lies.  /root/reference exists in
the authoring container only, so everything that needs this module either runs
there (tests/golden/make_golden.py, `-m "not gpu"` tests that skip when the
reference is absent) or reads the committed fixtures it produced.
"""
import importlib
import os
import sys

SHIM_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'tf_shim')
REFERENCE_ROOT = os.environ.get('DDSP_REFERENCE_ROOT', '/root/reference')


def available():
  return os.path.isfile(os.path.join(REFERENCE_ROOT, 'ddsp', 'core.py'))


def load():
  """Imports and returns the reference `ddsp` package (on the shim)."""
  if not available():
    raise RuntimeError('reference sources not found under %s' % REFERENCE_ROOT)
  try:
    import tensorflow as tf  # noqa: F401
    if 'numpy-shim' not in getattr(tf, '__version__', ''):
      raise RuntimeError('a real TensorFlow is importable; use it directly')
  except ImportError:
    sys.path.insert(0, SHIM_DIR)
  if REFERENCE_ROOT not in sys.path:
    sys.path.append(REFERENCE_ROOT)
  return importlib.import_module('ddsp')


def tf():
  load()
  return importlib.import_module('tensorflow')


def to_numpy(x):
  """numpy view of a shim tensor / nested dict of them."""
  if isinstance(x, dict):
    return {k: to_numpy(v) for k, v in x.items()}
  return x.numpy() if hasattr(x, 'numpy') else x
