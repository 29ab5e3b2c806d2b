"""Full-size (BASELINE.json shapes) GPU tests through size-independent
properties, plus oracle spot checks on a few items of the big batch.

C2: B=32, F=1000, K=100, nb=65, N=64000.  (C3's B=256 only changes the batch
axis, which the item-independence test covers.)
"""
import numpy as np
import pytest
import torch

from oracle import ddsp_oracle as oracle
from tests.util import rel_err, synth_inputs

import ddsp_b200
from ddsp_b200 import core

pytestmark = pytest.mark.gpu
B, F, K, NB, N = 32, 1000, 100, 65, 64000
TOL = 1e-4


@pytest.fixture(scope='module')
def c2():
  inp = synth_inputs(B, F, K, NB, N, seed=2024)
  dev = {k: torch.from_numpy(v).cuda() for k, v in inp.items()}
  amps, hd = core.harmonic_controls(dev['amps'], dev['harmonic_distribution'],
                                    dev['f0_hz'], 16000)
  mags = core.noise_controls(dev['noise_magnitudes'], -5.0)
  return inp, dev, amps, hd, mags


def _np(x):
  return x.detach().cpu().numpy()


def test_oracle_spot_check_full_length(c2):
  """Items 0 and 17 of the C2 batch against the float64 arbiter at N=64000."""
  inp, dev, amps, hd, mags = c2
  audio = _np(core.harmonic_synthesis(dev['f0_hz'], amps,
                                      harmonic_distribution=hd, n_samples=N))
  nz = _np(core.filtered_noise(mags, N, window_size=0, noise=dev['noise']))
  for b in (0, 17):
    sl = slice(b, b + 1)
    want = oracle.harmonic_synthesis(
        inp['f0_hz'][sl], _np(amps)[sl], harmonic_distribution=_np(hd)[sl],
        n_samples=N, dtype=np.float64)
    emax, el2 = rel_err(audio[sl], want)
    assert emax < TOL and el2 < TOL, ('harmonic', b, emax, el2)
    want_n = oracle.frequency_filter(inp['noise'][sl], _np(mags)[sl].astype(np.float64),
                                     window_size=0)
    emax, el2 = rel_err(nz[sl], want_n)
    assert emax < TOL and el2 < TOL, ('noise', b, emax, el2)


def test_constant_f0_phase_is_exact_at_full_length():
  """4 s of a 100-harmonic tone: inclusive-cumsum closed form, no drift."""
  f0 = np.full((2, F, 1), 77.7, np.float32)          # all 100 harmonics live
  f0[1] = 79.9
  amp = np.ones((2, F, 1), np.float32)
  hd = np.full((2, F, K), 1.0 / K, np.float32)
  got = _np(core.harmonic_synthesis(f0, amp, harmonic_distribution=hd, n_samples=N))
  t = (np.arange(N, dtype=np.float64) + 1) / 16000.0
  for b in range(2):
    fb = float(f0[b, 0, 0])                           # the float32 value
    k = np.arange(1, K + 1, dtype=np.float64)
    want = np.zeros(N)
    for kk in k:                                      # avoid an [N, K] temp
      want += np.sin(2 * np.pi * fb * kk * t) / K
    assert np.abs(got[b] - want).max() < 1e-4 * np.abs(want).max()


def test_harmonic_is_linear_in_amplitude(c2):
  _, dev, amps, hd, _ = c2
  a = core.harmonic_synthesis(dev['f0_hz'], amps, harmonic_distribution=hd, n_samples=N)
  b = core.harmonic_synthesis(dev['f0_hz'], amps * 2.0, harmonic_distribution=hd,
                              n_samples=N)
  emax, _ = rel_err(_np(b), 2.0 * _np(a))
  assert emax < 2e-6


def test_items_are_independent(c2):
  """Batch sharding: a shard computed alone equals its rows in the
  full batch - bit for bit (same kernel, same per-item arithmetic)."""
  _, dev, amps, hd, mags = c2
  full = core.harmonic_synthesis(dev['f0_hz'], amps, harmonic_distribution=hd, n_samples=N)
  part = core.harmonic_synthesis(dev['f0_hz'][8:16], amps[8:16],
                                 harmonic_distribution=hd[8:16], n_samples=N)
  assert torch.equal(full[8:16], part)
  fulln = core.filtered_noise(mags, N, window_size=0, noise=dev['noise'])
  partn = core.filtered_noise(mags[8:16], N, window_size=0, noise=dev['noise'][8:16])
  assert torch.equal(fulln[8:16], partn)


def test_harmonic_is_causal_in_frames(c2):
  """audio[t] depends on frames <= t // hop + 1 only: synthesising the first
  half of the frames reproduces the first half of the audio (minus one hop)."""
  _, dev, amps, hd, _ = c2
  full = core.harmonic_synthesis(dev['f0_hz'][:4], amps[:4],
                                 harmonic_distribution=hd[:4], n_samples=N)
  half = core.harmonic_synthesis(dev['f0_hz'][:4, :F // 2].contiguous(),
                                 amps[:4, :F // 2].contiguous(),
                                 harmonic_distribution=hd[:4, :F // 2].contiguous(),
                                 n_samples=N // 2)
  n_ok = N // 2 - N // F
  emax, _ = rel_err(_np(half[:, :n_ok]), _np(full[:, :n_ok]))
  assert emax < 1e-6


def test_noise_filter_is_linear_in_magnitudes(c2):
  _, dev, _, _, mags = c2
  m1, m2 = mags[:8], mags[8:16]
  nz = dev['noise'][:8]
  a = core.filtered_noise(m1, N, window_size=0, noise=nz)
  b = core.filtered_noise(m2, N, window_size=0, noise=nz)
  ab = core.filtered_noise(m1 + m2, N, window_size=0, noise=nz)
  emax, _ = rel_err(_np(ab), _np(a) + _np(b))
  assert emax < 5e-6
  a3 = core.filtered_noise(m1 * 3.0, N, window_size=0, noise=nz)
  emax, _ = rel_err(_np(a3), 3.0 * _np(a))
  assert emax < 5e-6


def test_flat_magnitudes_delay_by_two_samples(c2):
  """Flat magnitudes g: the 128-tap windowed IR is g * delta at tap 64 and the
  reference crops 62, so out[t] = g x[t-2] (SURVEY 8c known answer)."""
  _, dev, _, _, _ = c2
  g = 0.37
  mags = torch.full((4, F, NB), g, device='cuda')
  nz = dev['noise'][:4]
  out = _np(core.filtered_noise(mags, N, window_size=0, noise=nz))
  x = _np(nz)
  assert np.abs(out[:, 2:] - g * x[:, :-2]).max() < 2e-6
  assert np.abs(out[:, :2]).max() < 2e-6


def test_fused_decoder_equals_sum_of_parts(c2):
  """ProcessorGroup fused path (2 launches from raw outputs) == harmonic +
  noise computed through get_controls / get_signal, same Philox stream."""
  inp, dev, amps, hd, mags = c2
  harm = ddsp_b200.Harmonic(n_samples=N)
  noise = ddsp_b200.FilteredNoise(n_samples=N, window_size=0, seed=9)
  group = ddsp_b200.ProcessorGroup(dag=[
      (harm, ['amps', 'harmonic_distribution', 'f0_hz']),
      (noise, ['noise_magnitudes']),
      (ddsp_b200.Add(), ['filtered_noise/signal', 'harmonic/signal'])])
  feats = {k: dev[k] for k in ['amps', 'harmonic_distribution', 'f0_hz',
                               'noise_magnitudes']}
  fused = group(feats)                                    # Philox offset 0
  h = core.harmonic_synthesis(dev['f0_hz'], amps, harmonic_distribution=hd, n_samples=N)
  nzs = core.filtered_noise(mags, N, window_size=0, seed=9, offset=0)
  emax, el2 = rel_err(_np(fused), _np(h) + _np(nzs))
  assert emax < 1e-5 and el2 < 1e-5, (emax, el2)
  # successive calls draw fresh noise (synths.py:192-193 semantics)
  again = group(feats)
  assert not torch.equal(fused, again)


def test_in_kernel_noise_statistics():
  x = _np(core.uniform_noise(8, N, seed=5))
  assert x.min() >= -1.0 and x.max() < 1.0
  assert abs(x.mean()) < 5e-3 and abs(x.var() - 1.0 / 3.0) < 3e-3
  # neighbouring samples and neighbouring items are uncorrelated
  assert abs(np.mean(x[:, 1:] * x[:, :-1])) < 3e-3
  assert abs(np.mean(x[0] * x[1])) < 5e-3
