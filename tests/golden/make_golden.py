"""Writes tests/golden/*.npz: small input / output vectors of the hot path computed
by the float64 oracle (oracle/ddsp_oracle.py).

What these are NOT: outputs of the TensorFlow reference - TensorFlow cannot be
installed in this environment, so reference parity of the
oscillator numerics stays "unpinned" (DESIGN.md section 5).  What they are: a
regression anchor.  The oracle is the arbiter of every GPU parity test; if its
arithmetic drifts, `tests/test_oracle.py::test_oracle_matches_committed_golden`
fails on CPU before any kernel is blamed.

  python tests/golden/make_golden.py        # regenerate (only after a deliberate
                                            # oracle change; say why in the commit)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ddsp_oracle as o          # noqa: E402
from tests.util import synth_inputs          # noqa: E402


def c1_harmonic():
  """BASELINE.json configs[0]: Harmonic only, B=1, 16000 samples, 64 harmonics,
  250 frames."""
  inp = synth_inputs(1, 250, 64, 65, 16000, seed=101)
  ctl = o.harmonic_get_controls(inp['amps'], inp['harmonic_distribution'],
                                inp['f0_hz'], dtype=np.float32)
  audio = o.harmonic_synthesis(ctl['f0_hz'], ctl['amplitudes'],
                               harmonic_distribution=ctl['harmonic_distribution'],
                               n_samples=16000, dtype=np.float64)
  return dict(seed=101, audio=audio.astype(np.float32),
              amplitudes=ctl['amplitudes'], f0_hz=ctl['f0_hz'],
              hd_checksum=np.float64(ctl['harmonic_distribution'].astype(np.float64).sum()))


def decoder_small():
  """The ae.gin DAG from raw network outputs, B=2, F=25, N=1600, injected noise."""
  B, F, K, nb, N = 2, 25, 100, 65, 1600
  inp = synth_inputs(B, F, K, nb, N, seed=202)
  out = o.decoder(inp['amps'], inp['harmonic_distribution'], inp['f0_hz'],
                  inp['noise_magnitudes'], inp['noise'], n_samples=N, window_size=0,
                  dtype=np.float64)
  return dict(seed=202, audio=out['add']['signal'].astype(np.float32),
              harmonic=out['harmonic']['signal'].astype(np.float32),
              filtered_noise=out['filtered_noise']['signal'].astype(np.float32))


if __name__ == '__main__':
  np.savez_compressed(os.path.join(HERE, 'c1_harmonic.npz'), **c1_harmonic())
  np.savez_compressed(os.path.join(HERE, 'decoder_small.npz'), **decoder_small())
  print('wrote', sorted(f for f in os.listdir(HERE) if f.endswith('.npz')))
