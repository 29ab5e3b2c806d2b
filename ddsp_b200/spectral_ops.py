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
  window = _hann(frame_size, audio.device)
  return torch.fft.rfft(frames * window, n=fft_length, dim=-1)


def compute_mag(audio, size=2048, overlap=0.75, pad_end=True):
  """spectral_ops.compute_mag (spectral_ops.py:67-70)."""
  return torch.abs(stft(audio, frame_size=size, overlap=overlap, pad_end=pad_end))


# ---- CUDA pieces of the spectrogram loss (include/ddsp_b200.h) ---------------
_WINDOWS = {}


def _hann(frame_size, device):
  key = (int(frame_size), str(device))
  if key not in _WINDOWS:
    # tf.signal.hann_window: periodic for even lengths, symmetric for odd ones
    _WINDOWS[key] = torch.hann_window(int(frame_size),
                                      periodic=(int(frame_size) % 2 == 0),
                                      dtype=torch.float32, device=device)
  return _WINDOWS[key]


def _stream():
  return torch.cuda.current_stream().cuda_stream


class FrameWindowFn(torch.autograd.Function):
  """stft's framing + periodic Hann window (pad_end=True) as one CUDA kernel,
  and its transpose (windowed overlap-add of the frame gradients) for backward.
  audio [B, N] -> frames [B, ceil(N / step), frame_size]."""

  @staticmethod
  def forward(ctx, audio, frame_size, step):
    from ddsp_b200 import _lib
    audio = audio.contiguous()
    b, n = audio.shape
    n_frames = -(-n // step)
    frames = torch.empty((b, n_frames, frame_size), dtype=torch.float32,
                         device=audio.device)
    _lib.check(_lib.load().ddsp_b200_frame_window(
        audio.data_ptr(), _hann(frame_size, audio.device).data_ptr(),
        frames.data_ptr(), b, n, n_frames, frame_size, step, _stream()))
    ctx.meta = (b, n, n_frames, frame_size, step)
    return frames

  @staticmethod
  def backward(ctx, grad_frames):
    from ddsp_b200 import _lib
    b, n, n_frames, frame_size, step = ctx.meta
    grad_frames = grad_frames.contiguous()
    grad_audio = torch.empty((b, n), dtype=torch.float32, device=grad_frames.device)
    _lib.check(_lib.load().ddsp_b200_frame_window_adjoint(
        grad_frames.data_ptr(), _hann(frame_size, grad_frames.device).data_ptr(),
        grad_audio.data_ptr(), b, n, n_frames, frame_size, step, None, 0, _stream()))
    return grad_audio, None, None


def _frame_window(audio, frame_size, step):
  from ddsp_b200 import _lib
  audio = audio.contiguous()
  b, n = audio.shape
  n_frames = -(-n // step)
  frames = torch.empty((b, n_frames, frame_size), dtype=torch.float32,
                       device=audio.device)
  _lib.check(_lib.load().ddsp_b200_frame_window(
      audio.data_ptr(), _hann(frame_size, audio.device).data_ptr(),
      frames.data_ptr(), b, n, n_frames, frame_size, step, _stream()))
  return frames


class SpectralTermFn(torch.autograd.Function):
  """One FFT size of the 'L1' spectrogram loss (losses.py:102-127, 190-243):
  mag_weight * mean|mag_t - mag_v| + logmag_weight * mean|safe_log mag_t -
  safe_log mag_v| of audio [B, N] against the target's complex STFT.

  forward : framing + Hann (one kernel), cuFFT r2c, ONE pass over both STFTs that
            also leaves the gradient w.r.t. the value STFT, pre-scaled so that the
            transpose of rfft is a plain irfft.
  backward: cuFFT c2r, windowed overlap-add of the frame gradients (one kernel)."""

  @staticmethod
  def forward(ctx, stft_target, audio, frame_size, step, mag_weight, logmag_weight):
    from ddsp_b200 import _lib
    audio = audio.to(torch.float32).contiguous()
    frames = _frame_window(audio, frame_size, step)
    xv = torch.fft.rfft(frames, n=frame_size, dim=-1)
    del frames
    xt = stft_target.contiguous()
    if xt.shape != xv.shape:
      raise ValueError(f'target STFT {tuple(xt.shape)} vs value STFT {tuple(xv.shape)}')
    grad = torch.empty_like(xv)
    sums = torch.zeros(2, dtype=torch.float64, device=xv.device)
    m = xv.numel()
    _lib.check(_lib.load().ddsp_b200_spectral_l1(
        xt.data_ptr(), xv.data_ptr(), grad.data_ptr(), sums.data_ptr(), m,
        float(mag_weight), float(logmag_weight), xv.shape[-1], frame_size,
        _stream()))
    ctx.save_for_backward(grad)
    ctx.meta = (audio.shape[0], audio.shape[1], xv.shape[1], frame_size, step)
    w = torch.tensor([mag_weight / m, logmag_weight / m], dtype=torch.float64,
                     device=xv.device)
    return (sums * w).sum().to(torch.float32)

  @staticmethod
  def backward(ctx, grad_out):
    from ddsp_b200 import _lib
    (grad,) = ctx.saved_tensors
    b, n, n_frames, frame_size, step = ctx.meta
    grad_frames = torch.fft.irfft(grad, n=frame_size, dim=-1).contiguous()
    grad_audio = torch.empty((b, n), dtype=torch.float32, device=grad.device)
    _lib.check(_lib.load().ddsp_b200_frame_window_adjoint(
        grad_frames.data_ptr(), _hann(frame_size, grad.device).data_ptr(),
        grad_audio.data_ptr(), b, n, n_frames, frame_size, step, None, 0, _stream()))
    return None, grad_audio * grad_out, None, None, None, None


_LOSS_WEIGHTS = {}


class SpectralLossFn(torch.autograd.Function):
  """The whole multi-scale 'L1' spectrogram loss (losses.SpectralLoss.call,
  losses.py:194-243, `ae.gin` weights) as ONE autograd node: per FFT size framing +
  Hann (kernel), cuFFT r2c of target and value, one pass leaving both L1 sums and
  the value-STFT gradient IN PLACE of the value STFT; backward is cuFFT c2r plus a
  windowed overlap-add per size that accumulates, already scaled by the upstream
  gradient (read on the device), into a single dL/d audio buffer.  No elementwise
  torch op on either pass."""

  @staticmethod
  def forward(ctx, target, audio, fft_sizes, mag_weight, logmag_weight):
    from ddsp_b200 import _lib
    lib = _lib.load()
    audio = audio.to(torch.float32).contiguous()
    target = target.to(torch.float32).contiguous()
    b, n = audio.shape
    sums = torch.zeros((len(fft_sizes), 2), dtype=torch.float64, device=audio.device)
    grads, counts = [], []
    for idx, size in enumerate(fft_sizes):
      size = int(size)
      step = int(size * 0.25)
      xt = torch.fft.rfft(_frame_window(target, size, step), n=size, dim=-1)
      xv = torch.fft.rfft(_frame_window(audio, size, step), n=size, dim=-1)
      m = xv.numel()
      _lib.check(lib.ddsp_b200_spectral_l1(
          xt.data_ptr(), xv.data_ptr(), xv.data_ptr(), sums[idx].data_ptr(), m,
          float(mag_weight), float(logmag_weight), xv.shape[-1], -1, _stream()))
      del xt
      grads.append(xv)                   # now holds d loss_size / d X_value, irfft-ready
      counts.append(m)
    key = (tuple(counts), float(mag_weight), float(logmag_weight), str(audio.device))
    if key not in _LOSS_WEIGHTS:
      _LOSS_WEIGHTS[key] = torch.tensor(
          [[mag_weight / m, logmag_weight / m] for m in counts], dtype=torch.float64,
          device=audio.device)
    ctx.save_for_backward(*grads)
    ctx.meta = (b, n, tuple(int(sz) for sz in fft_sizes))
    return (sums * _LOSS_WEIGHTS[key]).sum().to(torch.float32)

  @staticmethod
  def backward(ctx, grad_out):
    from ddsp_b200 import _lib
    lib = _lib.load()
    b, n, sizes = ctx.meta
    grads = ctx.saved_tensors
    go = grad_out.to(torch.float32).contiguous()
    grad_audio = torch.empty((b, n), dtype=torch.float32, device=go.device)
    for idx, size in enumerate(sizes):
      step = int(size * 0.25)
      n_frames = -(-n // step)
      # unnormalised inverse (the 1/n pass would be an elementwise kernel over the
      # frames; spectral_l1 left the spectrum scaled for exactly this)
      gf = torch.fft.irfft(grads[idx], n=size, dim=-1, norm='forward').contiguous()
      _lib.check(lib.ddsp_b200_frame_window_adjoint(
          gf.data_ptr(), _hann(size, go.device).data_ptr(), grad_audio.data_ptr(), b, n,
          n_frames, size, step, go.data_ptr(), int(idx > 0), _stream()))
    return None, grad_audio, None, None, None


def stft_cuda(audio, frame_size, overlap=0.75):
  """stft(pad_end=True) for CUDA tensors through FrameWindowFn + cuFFT."""
  step = int(frame_size * (1.0 - overlap))
  fft_length = 1 << (int(frame_size) - 1).bit_length()
  frames = FrameWindowFn.apply(audio.to(torch.float32), int(frame_size), step)
  return torch.fft.rfft(frames, n=fft_length, dim=-1)
