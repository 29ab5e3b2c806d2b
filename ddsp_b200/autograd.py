"""Differentiable wrappers of the synthesis kernels (C4: decoder forward +
backward through SpectralLoss).

The reference differentiates through every TF op of `core.harmonic_synthesis` /
`core.frequency_filter`; here forward and backward are the hand-written CUDA
kernels, exposed as `torch.autograd.Function`s:

  * `HarmonicSynthesisFn`  - d amplitudes, d harmonic_distribution (d f0 is not
    built: in `ae.gin` f0 is data, `training/preprocessing.py:74-91`);
  * `FilteredNoiseFn`      - d magnitudes (the filter is linear in them).

`Harmonic.get_controls` / `FilteredNoise.get_controls` (exp_sigmoid, Nyquist
normalisation) are frame-rate and run as ordinary torch ops on this path, so
autograd carries the gradient on to the raw network outputs.
"""
import math

import torch

from ddsp_b200 import _lib
from ddsp_b200 import core


def _stream():
  return torch.cuda.current_stream().cuda_stream


class HarmonicSynthesisFn(torch.autograd.Function):
  """core.harmonic_synthesis (core.py:1048-1111), differentiable in
  amplitudes and harmonic_distribution."""

  @staticmethod
  def forward(ctx, f0_hz, amplitudes, harmonic_distribution, n_samples,
              sample_rate, amp_resample_method):
    f0_hz = core.torch_float32(f0_hz)
    amplitudes = core.torch_float32(amplitudes)
    harmonic_distribution = core.torch_float32(harmonic_distribution)
    ctx.save_for_backward(f0_hz, amplitudes, harmonic_distribution)
    ctx.cfg = (int(n_samples), float(sample_rate), amp_resample_method)
    return core.harmonic_synthesis(
        f0_hz, amplitudes, harmonic_distribution=harmonic_distribution,
        n_samples=n_samples, sample_rate=sample_rate,
        amp_resample_method=amp_resample_method)

  @staticmethod
  def backward(ctx, grad_audio):
    f0_hz, amplitudes, hd = ctx.saved_tensors
    n_samples, sample_rate, method = ctx.cfg
    if ctx.needs_input_grad[0]:
      raise NotImplementedError(
          'HarmonicSynthesisFn: f0_hz requires grad, but the gradient with respect '
          'to the fundamental frequency is not built (the reference gets it from '
          'TF autodiff through the phase cumsum).  Detach f0_hz, or differentiate '
          'core.oscillator_bank-style torch ops for that path.')
    b, f, k = hd.shape
    grad_audio = grad_audio.contiguous().to(torch.float32)
    g0 = torch.empty_like(hd)
    g1 = torch.empty_like(hd)
    _lib.check(_lib.load().ddsp_b200_harmonic_backward(
        f0_hz.data_ptr(), grad_audio.data_ptr(), g0.data_ptr(), g1.data_ptr(),
        b, f, k, n_samples, sample_rate, core.AMP_METHODS[method], _stream()))
    # dL/d(amp * hd)[i] = g0[i] + g1[i-1], frame F being a copy of frame F-1
    dha = g0
    dha[:, 1:] += g1[:, :-1]
    dha[:, -1] += g1[:, -1]
    d_hd = dha * amplitudes
    d_amp = (dha * hd).sum(-1, keepdim=True)
    return None, d_amp, d_hd, None, None, None


class FilteredNoiseFn(torch.autograd.Function):
  """FilteredNoise.get_signal (synths.py:181-196), differentiable in magnitudes."""

  @staticmethod
  def forward(ctx, magnitudes, n_samples, window_size, noise, seed, offset):
    magnitudes = core.torch_float32(magnitudes)
    ctx.cfg = (int(n_samples), int(window_size), int(seed), int(offset),
               tuple(magnitudes.shape))
    ctx.noise = None if noise is None else core.torch_float32(noise)
    return core.filtered_noise(magnitudes, n_samples, window_size=window_size,
                               noise=ctx.noise, seed=seed, offset=offset)

  @staticmethod
  def backward(ctx, grad_audio):
    n_samples, window_size, seed, offset, (b, f, nb) = ctx.cfg
    grad_audio = grad_audio.contiguous().to(torch.float32)
    dmags = torch.empty((b, f, nb), dtype=torch.float32, device=grad_audio.device)
    _lib.check(_lib.load().ddsp_b200_filtered_noise_backward(
        grad_audio.data_ptr(), 0 if ctx.noise is None else ctx.noise.data_ptr(),
        seed & (2**64 - 1), offset & (2**64 - 1), dmags.data_ptr(), b, f, nb,
        n_samples, window_size, _stream()))
    return dmags, None, None, None, None, None


def exp_sigmoid(x, exponent=10.0, max_value=2.0, threshold=1e-7):
  """core.exp_sigmoid (core.py:386-404) as differentiable torch ops."""
  return max_value * torch.sigmoid(x)**math.log(exponent) + threshold


def harmonic_controls(amps, harmonic_distribution, f0_hz, sample_rate=16000,
                      normalize_below_nyquist=True):
  """Harmonic.get_controls (synths.py:94-121) as differentiable torch ops."""
  amps = exp_sigmoid(amps)
  hd = exp_sigmoid(harmonic_distribution)
  if normalize_below_nyquist:
    k = hd.shape[-1]
    ratios = torch.linspace(1.0, float(k), k, device=hd.device, dtype=hd.dtype)
    hd = torch.where(f0_hz * ratios >= sample_rate / 2.0, torch.zeros_like(hd), hd)
  denom = hd.sum(-1, keepdim=True)
  denom = torch.where(denom == 0.0, torch.full_like(denom, 1e-7), denom)
  return amps, hd / denom


def decoder_train(amps, harmonic_distribution, f0_hz, noise_magnitudes,
                  n_samples=64000, sample_rate=16000, window_size=0,
                  initial_bias=-5.0, noise=None, seed=0, offset=0):
  """The `ae.gin` decoder (ae.gin:47-72) with gradients to amps,
  harmonic_distribution and noise_magnitudes: get_controls in torch, the two
  synthesizers as CUDA forward / backward kernels, Add in torch."""
  dev = core._device()
  amps = core.torch_float32(amps, dev) if not isinstance(amps, torch.Tensor) else amps
  a, h = harmonic_controls(amps, harmonic_distribution, f0_hz, sample_rate)
  harm = HarmonicSynthesisFn.apply(f0_hz, a, h, n_samples, sample_rate, 'window')
  mags = exp_sigmoid(noise_magnitudes + initial_bias)
  nz = FilteredNoiseFn.apply(mags, n_samples, window_size, noise, seed, offset)
  return harm + nz
