"""Effects processors that consume the decoder's audio: `Reverb` and `FIRFilter`
with the reference's constructors and semantics (`ddsp/effects.py:28-117,
283-325`; the next node after `Add` in `solo_instrument.gin:26-40`).

`Reverb` is a long linear time-invariant convolution (48000-tap impulse
response): `core.fft_convolve` routes impulse responses of 2048 taps and more to
the framed-FFT formulation on cuFFT; `FIRFilter` is the
time-varying filter of `FilteredNoise` applied to given audio and runs on the
hand-written IR + FIR kernels."""
import torch

from ddsp_b200 import core
from ddsp_b200 import processors


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
    audio = core.torch_float32(audio)
    if self.trainable:
      self.build(audio.device)
      ir = self._match_dimensions(audio, self._ir)
    elif ir is None:
      raise ValueError('Must provide "ir" tensor if Reverb trainable=False.')
    return {'audio': audio, 'ir': ir}

  def get_signal(self, audio, ir):
    """effects.py:103-117."""
    audio, ir = core.torch_float32(audio), core.torch_float32(ir)
    ir = self._mask_dry_ir(ir)
    wet = core.fft_convolve(audio, ir, padding='same', delay_compensation=0)
    return (wet + audio) if self._add_dry else wet


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
