"""Frequency scaling helpers and the Sinusoidal processor.

CPU: core.frequencies_sigmoid / unit_to_hz / hz_to_midi against the NumPy oracle
and the reference's own bounds test (synths_test.py:89-110).  GPU: the whole
processor (resample + oscillator_bank kernels) against the float64 oracle."""
import numpy as np
import pytest
import torch

from oracle import ddsp_oracle as o
from ddsp_b200 import core


@pytest.mark.parametrize('depth', [1, 4, 10])
def test_frequencies_sigmoid_matches_oracle(depth):
  rng = np.random.default_rng(depth)
  x = rng.normal(0, 3, (2, 7, 5 * depth)).astype(np.float32)
  got = core.frequencies_sigmoid(torch.from_numpy(x), depth=depth).numpy()
  want = o.frequencies_sigmoid(x, depth=depth)
  assert got.shape == want.shape == (2, 7, 5)
  np.testing.assert_allclose(got, want, rtol=2e-4, atol=2e-2)


@pytest.mark.parametrize('depth', [2, 16])
def test_frequencies_softmax_matches_oracle(depth):
  rng = np.random.default_rng(10 + depth)
  x = rng.normal(0, 2, (2, 5, 3 * depth)).astype(np.float32)
  got = core.frequencies_softmax(torch.from_numpy(x), depth=depth).numpy()
  want = o.frequencies_softmax(x, depth=depth)
  assert got.shape == want.shape == (2, 5, 3)
  np.testing.assert_allclose(got, want, rtol=2e-4, atol=2e-2)
  assert (got >= 20.0 - 1e-3).all() and (got <= 8000.0 + 1.0).all()


def test_frequencies_controls_are_bounded():
  """synths_test.py:89-110: 0 <= f <= 8000 Hz for any network output."""
  depth = 10
  x = torch.linspace(-100.0, 100.0, 100)[None, None, :, None].repeat(3, 10, 1, depth)
  f = core.frequencies_sigmoid(x, depth=depth, hz_min=0.0, hz_max=8000.0)
  assert tuple(f.shape) == (3, 10, 100)
  assert bool(((f >= 0.0) & (f <= 8000.0)).all())


def test_midi_hz_unit_round_trips():
  hz = torch.tensor([0.0, 27.5, 440.0, 8000.0])
  midi = core.hz_to_midi(hz)
  assert float(midi[0]) == 0.0 and abs(float(midi[2]) - 69.0) < 1e-4
  np.testing.assert_allclose(core.midi_to_hz(midi)[1:].numpy(), hz[1:].numpy(), rtol=1e-5)
  u = core.hz_to_unit(hz[1:], 20.0, 8000.0)
  np.testing.assert_allclose(core.unit_to_hz(u, 20.0, 8000.0).numpy(), hz[1:].numpy(),
                             rtol=1e-4)
  np.testing.assert_allclose(core.hz_to_midi(hz).numpy(), o.hz_to_midi(hz.numpy()),
                             rtol=1e-5, atol=1e-4)


@pytest.mark.gpu
@pytest.mark.parametrize('amp_method', ['window', 'linear'])
def test_sinusoidal_matches_oracle(amp_method):
  import ddsp_b200
  from tests.util import rel_err
  B, F, K, N = 2, 50, 12, 3200
  rng = np.random.default_rng(5)
  amps = rng.normal(0, 1, (B, F, K)).astype(np.float32)
  freqs = rng.normal(0, 2, (B, F, K)).astype(np.float32)
  synth = ddsp_b200.Sinusoidal(n_samples=N, sample_rate=16000,
                               amp_resample_method=amp_method)
  ctl = synth.get_controls(amps, freqs)
  want_ctl = o.sinusoidal_get_controls(amps, freqs)
  np.testing.assert_allclose(ctl['frequencies'].cpu().numpy(), want_ctl['frequencies'],
                             rtol=2e-4, atol=2e-2)
  np.testing.assert_allclose(ctl['amplitudes'].cpu().numpy(), want_ctl['amplitudes'],
                             rtol=2e-4, atol=1e-6)
  # signal from the SAME (float32) controls, so the comparison isolates get_signal
  a32 = ctl['amplitudes'].cpu().numpy()
  f32 = ctl['frequencies'].cpu().numpy()
  want = o.sinusoidal_get_signal(a32, f32, N, amp_resample_method=amp_method)
  got = synth.get_signal(**ctl).cpu().numpy()
  assert got.shape == (B, N)
  emax, el2 = rel_err(got, want)
  assert emax < 1e-4 and el2 < 1e-4, (emax, el2)
  # the Processor protocol: __call__ = get_signal(**get_controls(...))
  out = synth(amps, freqs, return_outputs_dict=True)
  assert set(out) == {'signal', 'controls'} and tuple(out['signal'].shape) == (B, N)


@pytest.mark.gpu
def test_sinusoidal_output_shape():
  """synths_test.py:75-87."""
  import ddsp_b200
  synth = ddsp_b200.Sinusoidal(n_samples=32000, sample_rate=16000)
  out = synth(np.zeros((3, 1000, 10), np.float32), np.zeros((3, 1000, 10), np.float32))
  assert tuple(out.shape) == (3, 32000)


# ---- ports of core_test.py:27-92 (UtilitiesTest); librosa's two formulas are
# hz = 440 * 2^((m - 69) / 12) and m = 12 (log2 hz - log2 440) + 69 ----------------
def test_midi_to_hz_is_accurate():
  midi = np.arange(128)
  want = 440.0 * 2.0 ** ((midi - 69.0) / 12.0)
  np.testing.assert_allclose(core.midi_to_hz(midi).numpy(), want, rtol=1e-5)


def test_hz_to_midi_is_accurate():
  hz = np.linspace(0.0, 20000.0, 128)
  with np.errstate(divide='ignore'):
    want = 12.0 * (np.log2(hz) - np.log2(440.0)) + 69.0
  want = np.where(hz <= 0.0, 0.0, want)
  np.testing.assert_allclose(core.hz_to_midi(hz).numpy(), want, rtol=1e-5, atol=1e-4)


@pytest.mark.parametrize('clip', [True, False])
def test_midi_unit_maps_are_accurate(clip):
  midi_min, midi_max = 20.0, 90.0
  midi = np.linspace(0.0, 127.0, 1000)
  want = (midi - midi_min) / (midi_max - midi_min)
  want = np.clip(want, 0.0, 1.0) if clip else want
  np.testing.assert_allclose(
      core.midi_to_unit(midi, midi_min=midi_min, midi_max=midi_max, clip=clip).numpy(),
      want, rtol=1e-5, atol=1e-6)
  unit = np.linspace(-1.0, 2.0, 1000)
  want = np.clip(unit, 0.0, 1.0) if clip else unit
  want = midi_min + (midi_max - midi_min) * want
  np.testing.assert_allclose(
      core.unit_to_midi(unit, midi_min=midi_min, midi_max=midi_max, clip=clip).numpy(),
      want, rtol=1e-5, atol=1e-4)


def test_unit_hz_maps_are_accurate():
  hz_min, hz_max = 20.0, 1000.0
  unit = np.linspace(0.0, 1.0, 128)
  hz = np.logspace(np.log10(hz_min), np.log10(hz_max), 128)
  np.testing.assert_allclose(core.unit_to_hz(unit, hz_min, hz_max).numpy(), hz, rtol=1e-4)
  np.testing.assert_allclose(core.hz_to_unit(hz, hz_min, hz_max).numpy(), unit,
                             rtol=1e-4, atol=1e-5)
