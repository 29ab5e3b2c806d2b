"""STFT magnitude helpers with the semantics of `ddsp/spectral_ops.py:34-70`.

Consumer of the decoder's audio in the C4 configuration (decoder -> multi-scale
SpectralLoss).  Built on torch's cuFFT-backed FFT as SURVEY.md section 8f-1
prescribes ("torch (cuFFT) first, not a hand kernel"); device-agnostic, so the
CPU tests can pin it against the NumPy oracle.
"""
import torch


def safe_log(x, eps=1e-5):
  """core.safe_log (core.py:213-216)."""
  return torch.log(torch.where(x <= 0.0, torch.full_like(x, eps), x))


def stft(audio, frame_size=2048, overlap=0.75, pad_end=True):
  """spectral_ops.stft (spectral_ops.py:34-47) = tf.signal.stft with a periodic
  Hann window, frame_step = frame_size * (1 - overlap), fft_length = enclosing
  power of two, and `pad_end` zero padding so that n_frames = ceil(N / step).
  Returns complex [batch, n_frames, fft_length // 2 + 1]."""
  if audio.dim() == 3:
    audio = audio.squeeze(-1)
  audio = audio.to(torch.float32)
  frame_size = int(frame_size)
  step = int(frame_size * (1.0 - overlap))
  n = audio.shape[-1]
  fft_length = 1 << (frame_size - 1).bit_length()
  if pad_end:
    n_frames = -(-n // step)
    padded = (n_frames - 1) * step + frame_size
    audio = torch.nn.functional.pad(audio, (0, max(0, padded - n)))
  else:
    n_frames = max(0, 1 + (n - frame_size) // step)
  frames = audio.unfold(-1, frame_size, step)[..., :n_frames, :]
  window = torch.hann_window(frame_size, periodic=True, dtype=torch.float32,
                             device=audio.device)
  return torch.fft.rfft(frames * window, n=fft_length, dim=-1)


def compute_mag(audio, size=2048, overlap=0.75, pad_end=True):
  """spectral_ops.compute_mag (spectral_ops.py:67-70)."""
  return torch.abs(stft(audio, frame_size=size, overlap=overlap, pad_end=pad_end))
