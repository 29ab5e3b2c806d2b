"""Effects processors that consume the decoder's audio: `Reverb`,
`FilteredNoiseReverb` and `FIRFilter` with the reference's constructors and
semantics (`ddsp/effects.py:28-117, 202-278, 283-325`; the next
node after `Add` in `solo_instrument.gin:26-40`).

`Reverb` is a long linear time-invariant convolution (48000-tap impulse
response): `core.fft_convolve` routes one impulse response of 2048 taps and more
per item to the hand-written partitioned overlap-save convolution
(csrc/longconv.cuh); `FilteredNoiseReverb` draws that impulse response from a
`FilteredNoise` synthesizer; `FIRFilter` is the time-varying filter of
`FilteredNoise` applied to given audio and runs on the IR + FIR kernels."""
import torch

from ddsp_b200 import core
from ddsp_b200 import processors
from ddsp_b200 import synths


class Reverb(processors.Processor):
  """Convolutional (FIR) reverb (effects.py:28-117)."""

  def __init__(self, trainable=False, reverb_length=48000, add_dry=True,
               name='reverb'):
    super().__init__(name=name, trainable=trainable)
    self._reverb_length = reverb_length
    self._add_dry = add_dry
    self._ir = None

  def _mask_dry_ir(self, ir):
    """effects.py:50-59: zero the first tap (the dry path)."""
    if ir.dim() == 1:
      ir = ir[None, :]
    if ir.dim() == 3:
      ir = ir[:, :, 0]
    dry_mask = torch.zeros((ir.shape[0], 1), dtype=torch.float32, device=ir.device)
    return torch.cat([dry_mask, ir[:, 1:]], dim=1)

  def _match_dimensions(self, audio, ir):
    """effects.py:61-68."""
    if ir.dim() == 1:
      ir = ir[None, :]
    return ir.repeat(int(audio.shape[0]), 1)

  def build(self, device=None):
    """effects.py:70-79: the single learned impulse response, N(0, 1e-6)."""
    if self.trainable and self._ir is None:
      self._ir = (1e-6 * torch.randn(self._reverb_length, dtype=torch.float32,
                                     device=device)).requires_grad_(True)

  def get_controls(self, audio, ir=None):
    """effects.py:81-101."""
    if not self.trainable and ir is None:          # before any device work
      raise ValueError('Must provide "ir" tensor if Reverb trainable=False.')
    audio = core.torch_float32(audio)
    if self.trainable:
      self.build(audio.device)
      ir = self._match_dimensions(audio, self._ir)
    return {'audio': audio, 'ir': ir}

  def get_signal(self, audio, ir):
    """effects.py:103-117."""
    audio, ir = core.torch_float32(audio), core.torch_float32(ir)
    ir = self._mask_dry_ir(ir)
    wet = core.fft_convolve(audio, ir, padding='same', delay_compensation=0)
    return (wet + audio) if self._add_dry else wet


class FilteredNoiseReverb(Reverb):
  """Impulse response = the output of a filtered-noise synthesizer
  (effects.py:202-278): `get_controls` runs `FilteredNoise(n_samples =
  reverb_length)` on the magnitudes (given per item, or ONE learned
  [n_frames, n_filter_banks] set tiled over the batch when trainable), `get_signal`
  is `Reverb`'s."""

  def __init__(self, trainable=False, reverb_length=48000, window_size=257,
               n_frames=1000, n_filter_banks=16, scale_fn=core.exp_sigmoid,
               initial_bias=-3.0, add_dry=True, name='filtered_noise_reverb'):
    super().__init__(name=name, add_dry=add_dry, trainable=trainable)
    self._n_frames = n_frames
    self._n_filter_banks = n_filter_banks
    self._synth = synths.FilteredNoise(n_samples=reverb_length,
                                       window_size=window_size,
                                       scale_fn=scale_fn,
                                       initial_bias=initial_bias)
    self._magnitudes = None

  def build(self, device=None):
    """effects.py:240-249: the learned magnitudes, N(0, 1e-2)."""
    if self.trainable and self._magnitudes is None:
      self._magnitudes = (1e-2 * torch.randn(
          self._n_frames, self._n_filter_banks, dtype=torch.float32,
          device=device)).requires_grad_(True)

  def _synth_ir(self, magnitudes):
    """`self._synth(magnitudes)` (effects.py:272); with gradients to the magnitudes
    when they ask for them (the synthesizer's autograd node, exp_sigmoid as
    differentiable torch ops on the [n_frames, n_filter_banks] controls)."""
    if (isinstance(magnitudes, torch.Tensor) and magnitudes.requires_grad and
        torch.is_grad_enabled()):
      from ddsp_b200 import autograd as _ag
      syn = self._synth
      if syn.scale_fn is core.exp_sigmoid:
        mags = _ag.exp_sigmoid(magnitudes + syn.initial_bias)
      elif syn.scale_fn is not None:
        mags = syn.scale_fn(magnitudes + syn.initial_bias)
      else:
        mags = magnitudes
      return _ag.FilteredNoiseFn.apply(mags, syn.n_samples, syn.window_size,
                                       syn.injected_noise, syn.seed,
                                       syn.next_offset())
    return self._synth(magnitudes)

  def get_controls(self, audio, magnitudes=None):
    """effects.py:251-277."""
    if not self.trainable and magnitudes is None:  # before any device work
      raise ValueError('Must provide "magnitudes" tensor if '
                       'FilteredNoiseReverb trainable=False.')
    audio = core.torch_float32(audio)
    if self.trainable:
      self.build(audio.device)
      magnitudes = self._magnitudes[None, :]
    ir = self._synth_ir(magnitudes)
    if self.trainable:
      ir = self._match_dimensions(audio, ir)
    return {'audio': audio, 'ir': ir}


class FIRFilter(processors.Processor):
  """Linear time-varying FIR filter (effects.py:283-325)."""

  def __init__(self, window_size=257, scale_fn=core.exp_sigmoid, name='fir_filter'):
    super().__init__(name=name)
    self.window_size = window_size
    self.scale_fn = scale_fn

  def get_controls(self, audio, magnitudes):
    if self.scale_fn is not None:
      magnitudes = self.scale_fn(core.torch_float32(magnitudes))
    return {'audio': audio, 'magnitudes': magnitudes}

  def get_signal(self, audio, magnitudes):
    return core.frequency_filter(audio, magnitudes, window_size=self.window_size)
