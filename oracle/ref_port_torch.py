"""Timed CPU baseline: op-by-op float32 port of the reference decoder path.

TEST / BENCH INFRASTRUCTURE ONLY (see oracle/ddsp_oracle.py header).  This is
"the reference's own CPU implementation" as far as it can exist here: the
reference runs on TensorFlow, which is not installable in this image, so each
TF op of ddsp/core.py + ddsp/synths.py is replaced by the torch-CPU op with the
same semantics, materialising the same [B, N, K] float32 intermediates TF does
and using all host threads torch is given (TF-CPU/Eigen also parallelises per
op).  Labelled kind="port" in bench.py - "restated reference, not TF".

Checked against oracle/ddsp_oracle.py (float32 mode) in tests/test_oracle.py.
"""
import math

import numpy as np
import torch

F32 = torch.float32


def _tf_hann(n):
  """tf.signal.hann_window(n): periodic for even n, symmetric for odd n
  (window_ops._raised_cosine_window's `even` term)."""
  return torch.hann_window(n, periodic=(n % 2 == 0), dtype=F32)


def exp_sigmoid(x):
  """core.py:386-404."""
  return 2.0 * torch.sigmoid(x)**math.log(10.0) + 1e-7


def _resample_linear(x, n_out):
  """core.py:613-627: tf.compat.v1.image.resize BILINEAR, align_corners=False."""
  n_in = x.shape[1]
  scale = np.float32(n_in) / np.float32(n_out)
  src = torch.arange(n_out, dtype=F32) * float(scale)
  lo = torch.floor(src)
  frac = (src - lo)[None, :, None]
  lo = lo.long()
  hi = torch.clamp(torch.ceil(src).long(), max=n_in - 1)
  top = x[:, lo, :]
  return top + (x[:, hi, :] - top) * frac


def _upsample_with_windows(x, n_out):
  """core.py:645-714 - literal windowed overlap-add (incl. both transposes)."""
  x = torch.cat([x, x[:, -1:, :]], dim=1)
  n_frames = x.shape[1]
  hop = n_out // (n_frames - 1)
  window = _tf_hann(2 * hop)
  x = x.permute(0, 2, 1)                                  # [B, C, frames]
  xw = x[:, :, :, None] * window[None, None, None, :]     # [B, C, frames, 2hop]
  b, c, f, w = xw.shape
  # overlap_and_add with 50 % overlap: two interleaved halves
  out = torch.zeros((b, c, (f + 1) * hop), dtype=F32)
  out[:, :, :f * hop] += xw[..., :hop].reshape(b, c, f * hop)
  out[:, :, hop:] += xw[..., hop:].reshape(b, c, f * hop)
  out = out.permute(0, 2, 1)
  return out[:, hop:-hop, :]


def harmonic_controls(amps, hd, f0, sample_rate=16000):
  """synths.py:94-121."""
  amps = exp_sigmoid(amps)
  hd = exp_sigmoid(hd)
  k = hd.shape[-1]
  hf = f0 * torch.linspace(1.0, float(k), k)[None, None, :]
  hd = torch.where(hf >= sample_rate / 2.0, torch.zeros_like(hd), hd)
  denom = hd.sum(-1, keepdim=True)
  denom = torch.where(denom == 0.0, torch.full_like(denom, 1e-7), denom)
  return amps, hd / denom


def harmonic_signal(amps, hd, f0, n_samples, sample_rate=16000):
  """synths.py:123-146 -> core.py:1048-1111, 911-962 (tf.cumsum variant)."""
  k = hd.shape[-1]
  hf = f0 * torch.linspace(1.0, float(k), k)[None, None, :]
  ha = amps * hd
  fe = _resample_linear(hf, n_samples)
  ae = _upsample_with_windows(ha, n_samples)
  ae = torch.where(fe >= sample_rate / 2.0, torch.zeros_like(ae), ae)
  omegas = fe * (2.0 * math.pi)
  omegas = omegas / float(sample_rate)
  phases = torch.cumsum(omegas, dim=1)
  return (ae * torch.sin(phases)).sum(-1)


def noise_controls(mags, initial_bias=-5.0):
  """synths.py:165-179."""
  return exp_sigmoid(mags + initial_bias)


def noise_signal(mags, n_samples, window_size=0, noise=None):
  """synths.py:181-196 -> core.py:1628-1655, 1534-1565, 1477-1531, 1382-1473."""
  b, f, nb = mags.shape
  if noise is None:
    noise = torch.rand((b, n_samples), dtype=F32) * 2.0 - 1.0
  ir = torch.fft.irfft(torch.complex(mags, torch.zeros_like(mags)))
  ir_size = ir.shape[-1]
  ws = window_size if 0 < window_size <= ir_size else ir_size
  window = _tf_hann(ws)
  padding = ir_size - ws
  if padding > 0:
    half = (ws + 1) // 2
    window = torch.cat([window[half:], torch.zeros(padding), window[:half]])
  else:
    window = torch.fft.fftshift(window)
  ir = window * ir
  if padding > 0:
    ir = torch.cat([ir[..., ir_size - (half - 1) + 1:], ir[..., :half + 1]], -1)
  else:
    ir = torch.fft.fftshift(ir, dim=-1)
  s = ir.shape[-1]
  frame = int(np.ceil(n_samples / f))
  pad = f * frame - n_samples
  frames = torch.nn.functional.pad(noise, (0, pad)).reshape(b, f, frame)
  fft_size = int(2**np.ceil(np.log2(s + frame - 1)))
  out = torch.fft.irfft(torch.fft.rfft(frames, fft_size) *
                        torch.fft.rfft(ir, fft_size), fft_size)
  # overlap_and_add(hop=frame)
  total = (f - 1) * frame + fft_size
  y = torch.nn.functional.fold(
      out.transpose(1, 2), output_size=(1, total), kernel_size=(1, fft_size),
      stride=(1, frame)).reshape(b, total)
  start = (s - 1) // 2 - 1
  return y[:, start:start + n_samples]


def decoder(amps, hd, f0, mags, n_samples=64000, sample_rate=16000,
            window_size=0, noise=None):
  """ae.gin:47-72 DAG: Harmonic -> FilteredNoise -> Add."""
  a, h = harmonic_controls(amps, hd, f0, sample_rate)
  harm = harmonic_signal(a, h, f0, n_samples, sample_rate)
  nz = noise_signal(noise_controls(mags), n_samples, window_size, noise)
  return nz + harm


def spectral_loss(target, audio, fft_sizes=(2048, 1024, 512, 256, 128, 64),
                  mag_weight=1.0, logmag_weight=1.0):
  """losses.SpectralLoss.call with loss_type='L1' (losses.py:194-243) on
  spectral_ops.compute_mag (spectral_ops.py:34-70): tf.signal.stft(pad_end=True),
  periodic Hann, frame_step = size / 4 - op by op in float32 torch-CPU ops."""
  loss = 0.0
  n = audio.shape[-1]
  for size in fft_sizes:
    step = size // 4
    n_frames = -(-n // step)
    pad = (n_frames - 1) * step + size - n
    win = _tf_hann(size)

    def mag(x):
      fr = torch.nn.functional.pad(x, (0, pad)).unfold(-1, size, step)
      return torch.abs(torch.fft.rfft(fr * win, dim=-1))

    t, v = mag(target), mag(audio)
    if mag_weight > 0:
      loss = loss + mag_weight * torch.mean(torch.abs(t - v))
    if logmag_weight > 0:
      slog = lambda m: torch.log(torch.where(m <= 0.0, torch.full_like(m, 1e-5), m))
      loss = loss + logmag_weight * torch.mean(torch.abs(slog(t) - slog(v)))
  return loss


def train_step(amps, hd, f0, mags, target, noise=None, n_samples=64000):
  """configs[3] on the CPU: decoder forward from raw network outputs, multi-scale
  SpectralLoss, backward to the three raw inputs (torch autograd standing in for
  TF's).  Returns the loss value."""
  leaves = [t.detach().clone().requires_grad_(True) for t in (amps, hd, mags)]
  audio = decoder(leaves[0], leaves[1], f0, leaves[2], n_samples=n_samples, noise=noise)
  loss = spectral_loss(target, audio)
  loss.backward()
  return float(loss)
