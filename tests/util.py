"""Shared synthetic inputs for tests and bench."""
import numpy as np


def synth_inputs(B, F, K, nb, N, seed=1234, sample_rate=16000, f0_lo=80.0,
                 f0_hi=800.0):
  """Raw network-output-like inputs of the decoder, float32 numpy.

  f0: per-item base U[f0_lo, f0_hi] Hz with a 3 % / 5 Hz vibrato, clipped to
  [20, 2000]; amps / harmonic_distribution / noise magnitudes ~ N(0,1) (the
  pre-get_controls distribution of processors_test.py:35-42); noise U[-1,1).
  """
  rng = np.random.default_rng(seed)
  hop = N / F
  t_frame = np.arange(F) * hop / sample_rate
  base = rng.uniform(f0_lo, f0_hi, size=(B, 1))
  ph = rng.uniform(0, 2 * np.pi, size=(B, 1))
  f0 = base * (1.0 + 0.03 * np.sin(2 * np.pi * 5.0 * t_frame[None, :] + ph))
  f0 = np.clip(f0, 20.0, 2000.0)[..., None].astype(np.float32)
  return {
      'f0_hz': f0,
      'amps': rng.standard_normal((B, F, 1)).astype(np.float32),
      'harmonic_distribution': rng.standard_normal((B, F, K)).astype(np.float32),
      'noise_magnitudes': rng.standard_normal((B, F, nb)).astype(np.float32),
      'noise': rng.uniform(-1.0, 1.0, size=(B, N)).astype(np.float32),
  }


def rel_err(a, b):
  """(max-abs error / max-abs reference, relative L2 error)."""
  a = np.asarray(a, np.float64)
  b = np.asarray(b, np.float64)
  peak = max(np.abs(b).max(), 1e-30)
  l2 = np.sqrt(((a - b)**2).sum() / max((b**2).sum(), 1e-60))
  return np.abs(a - b).max() / peak, l2
