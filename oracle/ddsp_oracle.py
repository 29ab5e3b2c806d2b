"""CPU oracle for the DDSP Harmonic + FilteredNoise decoder path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import it.  The product path (`ddsp_b200/`) never does, and fails
loudly when the CUDA library is missing.

What it is: an op-by-op NumPy restatement of the reference algorithm
(magenta/ddsp @ 88621d2, v3.7.0), every function citing the reference
file:line it follows.  It travels to the GPU box (where /root/reference does
not exist) and is what the `-m gpu` parity tests call.

Pinning status (round 2):
  * PINNED TO THE REFERENCE ITSELF.  The unmodified reference package
    (/root/reference/ddsp) runs in the authoring container on a NumPy stand-in
    for its TensorFlow primitives (oracle/tf_shim, loader oracle/ref_on_shim.py;
    the reference's own core_test / synths_test / processors_test pass on it,
    96 of 96).  tests/test_reference_pin.py checks, whenever /root/reference is
    present, that this file's float32 mode equals the reference run in float32
    to <= 2e-7 and its float64 mode equals the reference run "wide" (the same
    reference code evaluated in double precision) to <= 1e-9 - on configs[0], a
    configs[1] item, the decoder DAG, both `use_angular_cumsum` values,
    harmonic_shifts, every resample method, the SpectralLoss value.
    tests/golden/*.npz are outputs OF THE REFERENCE produced that way
    (tests/golden/make_golden.py) and are compared against both this file (CPU)
    and the CUDA path (GPU) wherever the tests run.
  * What stays third-party: the TF primitives themselves (`tf.compat.v1.image.
    resize`, `tf.signal.*`, `tf.cumsum`; setup.py:57 `tensorflow<=2.11`, absent
    from /root/reference and not installable here) are restated from their
    published semantics in BOTH this file and the shim; the reference's real
    tests anchor them (core_test.py:242-267 resample incl. 'cubic', 484-503,
    730-757 fft_convolve vs scipy, 759-785, 825-886).  The last ulp of Eigen's
    elementary functions and reduction order is not reproducible with NumPy.

Two arithmetic modes, selected by `dtype`:
  * np.float64 - the arbiter for the <=1e-4 relative gate ("exact" maths of the
    reference formulae; index math for the bilinear resize follows TF's float32
    scale computation only when `tf_index_math=True`).
  * np.float32 - "TF order": every intermediate rounded to float32, cumsum
    accumulated sequentially in float32 as TF-CPU does.  Used to report the
    reference's own phase-accumulation envelope, and as the timed CPU baseline.
"""
from __future__ import annotations

import numpy as np

TWO_PI = 2.0 * np.pi


# ----------------------------------------------------------------------------
# Small helpers (core.py:31-36, 207-216, 386-404)
# ----------------------------------------------------------------------------
def _as(x, dtype):
  return np.asarray(x, dtype=dtype)


def safe_divide(numerator, denominator, eps=1e-7):
  """core.py:207-210 - 0 denominators become eps."""
  dtype = np.result_type(numerator, denominator)
  safe = np.where(denominator == 0.0, np.asarray(eps, dtype), denominator)
  return numerator / safe


def safe_log(x, eps=1e-5):
  """core.py:213-216."""
  x = np.asarray(x)
  return np.log(np.where(x <= 0.0, np.asarray(eps, x.dtype), x))


def exp_sigmoid(x, exponent=10.0, max_value=2.0, threshold=1e-7,
                dtype=np.float64):
  """core.py:386-404: max_value * sigmoid(x)**log(exponent) + threshold."""
  x = _as(x, dtype)
  sig = 1.0 / (1.0 + np.exp(-x))
  sig = sig.astype(dtype)
  out = (np.asarray(max_value, dtype) *
         np.power(sig, np.asarray(np.log(exponent), dtype)) +
         np.asarray(threshold, dtype))
  return out.astype(dtype)


# ----------------------------------------------------------------------------
# TF op restatements
# ----------------------------------------------------------------------------
def hann_window(n, dtype=np.float64):
  """tf.signal.hann_window(n) (periodic=True): 0.5 - 0.5 cos(2 pi k / d) with d = n
  for even n and d = n - 1 for odd n - TensorFlow's `_raised_cosine_window`
  (window_ops.py: `n = window_length + periodic * even - 1`) gives the SYMMETRIC
  window whenever the length is odd, `periodic` or not.  (Until the last session of
  round 2 this function used d = n throughout; the unmodified reference on the shim
  exposed it on odd filter windows, e.g. window_size=257 with more than 129 bins.)"""
  if n == 1:
    return np.ones([1], dtype)
  k = np.arange(n, dtype=np.float64)
  d = n if n % 2 == 0 else n - 1
  return (0.5 - 0.5 * np.cos(TWO_PI * k / d)).astype(dtype)


def overlap_and_add(frames, hop):
  """tf.signal.overlap_and_add: [..., n_frames, frame_len] -> [..., out_len]."""
  n_frames, frame_len = frames.shape[-2], frames.shape[-1]
  out_len = (n_frames - 1) * hop + frame_len
  out = np.zeros(frames.shape[:-2] + (out_len,), frames.dtype)
  for i in range(n_frames):
    out[..., i * hop:i * hop + frame_len] += frames[..., i, :]
  return out


def frame_pad_end(x, frame_size, hop):
  """tf.signal.frame(x, frame_size, hop, pad_end=True) on the last axis."""
  n = x.shape[-1]
  n_frames = -(-n // hop)
  padded_len = (n_frames - 1) * hop + frame_size
  pad = padded_len - n
  if pad > 0:
    x = np.concatenate([x, np.zeros(x.shape[:-1] + (pad,), x.dtype)], axis=-1)
  idx = np.arange(n_frames)[:, None] * hop + np.arange(frame_size)[None, :]
  return x[..., idx]


def _bilinear_indices(n_in, n_out, align_corners, tf_index_math):
  """Index math of tf.compat.v1.image.resize(BILINEAR) along one axis.

  TF (resize_bilinear_op / image_resizer_state.h, legacy scaler) computes
  scale = in/out (or (in-1)/(out-1) when align_corners) in float32,
  src = out_idx * scale in float32, lower = floor(src),
  upper = min(ceil(src), in-1), lerp = src - floor(src).
  With tf_index_math=False the same quantities are computed exactly (float64),
  which coincides with TF whenever in/out is exactly representable (hop a
  power of two - every reference config).
  """
  if align_corners and n_out > 1:
    num, den = n_in - 1, n_out - 1
  else:
    num, den = n_in, n_out
  t = np.arange(n_out)
  if tf_index_math:
    scale = np.float32(num) / np.float32(den)
    src = (t.astype(np.float32) * scale).astype(np.float32)
    lo = np.floor(src)
    frac = (src - lo).astype(np.float32).astype(np.float64)
    hi = np.minimum(np.ceil(src), n_in - 1)
  else:
    lo = (t * num) // den
    frac = ((t * num) % den) / float(den)
    hi = np.minimum(lo + (frac > 0), n_in - 1)
  lo = np.maximum(lo, 0).astype(np.int64)
  hi = np.asarray(hi).astype(np.int64)
  return lo, hi, frac


def resize_bilinear_v1(x, n_out, align_corners=False, tf_index_math=False):
  """tf.compat.v1.image.resize(BILINEAR, align_corners) on axis 1 of [B,T,C].

  core.py:613-621 (`_image_resize`).  out = x[lo] + (x[hi] - x[lo]) * lerp.
  """
  dtype = x.dtype
  lo, hi, frac = _bilinear_indices(x.shape[1], n_out, align_corners,
                                   tf_index_math)
  frac = frac.astype(dtype)[None, :, None]
  top = x[:, lo, :]
  bottom = x[:, hi, :]
  return (top + ((bottom - top).astype(dtype) * frac).astype(dtype)).astype(dtype)


def resize_nearest_v1(x, n_out, align_corners=False):
  """tf.compat.v1.image.resize(NEAREST_NEIGHBOR) on axis 1 (legacy scaler)."""
  n_in = x.shape[1]
  if align_corners and n_out > 1:
    scale = np.float32(n_in - 1) / np.float32(n_out - 1)
    # C roundf: halves round away from zero (resize_nearest_neighbor_op.cc)
    src = np.floor(np.arange(n_out, dtype=np.float32) * scale + np.float32(0.5))
  else:
    scale = np.float32(n_in) / np.float32(n_out)
    src = np.floor(np.arange(n_out, dtype=np.float32) * scale)
  idx = np.minimum(src.astype(np.int64), n_in - 1)
  return x[:, idx, :]


_CUBIC_TABLE_SIZE = 1 << 10
_CUBIC_TABLE = []


def _cubic_coeffs_table():
  """Weights of TensorFlow's legacy bicubic kernel (tensorflow/core/kernels/image/
  resize_bicubic_op.cc, InitCoeffsTable with A = -0.75, half_pixel_centers =
  false - the kernel `tf.compat.v1.image.resize(BICUBIC)` runs, core.py:629): the
  Keys cubic sampled at 1025 offsets and stored as float32.  TensorFlow is a
  third-party dependency absent from /root/reference (setup.py:57 pins
  `tensorflow<=2.11`); this restates its published algorithm."""
  if not _CUBIC_TABLE:
    a = -0.75
    tab = np.empty((_CUBIC_TABLE_SIZE + 1, 2), np.float32)
    for i in range(_CUBIC_TABLE_SIZE + 1):
      x = i * 1.0 / _CUBIC_TABLE_SIZE
      tab[i, 0] = ((a + 2) * x - (a + 3)) * x * x + 1
      x += 1.0
      tab[i, 1] = ((a * x - 5 * a) * x + 8 * a) * x - 4 * a
    _CUBIC_TABLE.append(tab)
  return _CUBIC_TABLE[0]


def resize_bicubic_v1(x, n_out, align_corners=False):
  """tf.compat.v1.image.resize(BICUBIC, align_corners) on axis 1 of [B,T,C]
  (legacy scaler, GetWeightsAndIndices): in = out_idx * scale (float32),
  loc = floor(in), offset = lrintf((in - loc) * 1024), taps loc-1 .. loc+2 clamped
  to the ends, weights table[offset] / table[1024 - offset]."""
  dtype = x.dtype
  n_in = x.shape[1]
  tab = _cubic_coeffs_table()
  if align_corners and n_out > 1:
    scale = np.float32(n_in - 1) / np.float32(n_out - 1)
  else:
    scale = np.float32(n_in) / np.float32(n_out)
  src = (np.arange(n_out, dtype=np.float32) * scale).astype(np.float32)
  loc = np.floor(src).astype(np.int64)
  delta = (src - loc.astype(np.float32)).astype(np.float32)
  off = np.rint(delta * np.float32(_CUBIC_TABLE_SIZE)).astype(np.int64)
  w = [tab[off, 1], tab[off, 0], tab[_CUBIC_TABLE_SIZE - off, 0],
       tab[_CUBIC_TABLE_SIZE - off, 1]]
  out = np.zeros((x.shape[0], n_out, x.shape[2]), dtype)
  for j in range(4):
    idx = np.clip(loc - 1 + j, 0, n_in - 1)
    out = (out + (x[:, idx, :] * w[j].astype(dtype)[None, :, None]).astype(dtype)
           ).astype(dtype)
  return out


# ----------------------------------------------------------------------------
# Resampling (core.py:573-714)
# ----------------------------------------------------------------------------
def upsample_with_windows(inputs, n_timesteps, add_endpoint=True,
                          dtype=np.float64):
  """core.py:645-714 - literal Hann overlap-add upsampler."""
  inputs = _as(inputs, dtype)
  if inputs.ndim != 3:
    raise ValueError('Upsample_with_windows() only supports 3 dimensions, '
                     'not {}.'.format(inputs.shape))
  if add_endpoint:
    inputs = np.concatenate([inputs, inputs[:, -1:, :]], axis=1)
  n_frames = int(inputs.shape[1])
  n_intervals = n_frames - 1
  if n_frames >= n_timesteps:
    raise ValueError('Upsample with windows cannot be used for downsampling'
                     'More input frames ({}) than output timesteps ({})'.format(
                         n_frames, n_timesteps))
  if n_timesteps % n_intervals != 0.0:
    minus_one = '' if add_endpoint else ' - 1'
    raise ValueError(
        'For upsampling, the target the number of timesteps must be divisible '
        'by the number of input frames{}. (timesteps:{}, frames:{}, '
        'add_endpoint={}).'.format(minus_one, n_timesteps, n_frames,
                                   add_endpoint))
  hop_size = n_timesteps // n_intervals
  window = hann_window(2 * hop_size, dtype)
  x = np.transpose(inputs, [0, 2, 1])            # [B, C, frames]
  x_windowed = (x[:, :, :, None] * window[None, None, None, :]).astype(dtype)
  x = overlap_and_add(x_windowed, hop_size)      # [B, C, (frames+1)*hop]
  x = np.transpose(x, [0, 2, 1])
  return x[:, hop_size:-hop_size, :]


def resample(inputs, n_timesteps, method='linear', add_endpoint=True,
             dtype=np.float64, tf_index_math=False):
  """core.py:573-642."""
  inputs = _as(inputs, dtype)
  is_1d = inputs.ndim == 1
  is_2d = inputs.ndim == 2
  is_4d = inputs.ndim == 4
  shape_4d = inputs.shape
  if is_4d:
    # core.py:616-621: the image is resized to [n_timesteps, shape[2]] - the
    # n_freq axis keeps its size (an identity for every kernel), so a 4-D input
    # is the 3-D case with n_freq * channels channels.  'window' only takes 3-D.
    if method == 'window':
      raise ValueError('Upsample_with_windows() only supports 3 dimensions, '
                       'not {}.'.format(inputs.shape))
    inputs = inputs.reshape(shape_4d[0], shape_4d[1], -1)
  if is_1d:
    inputs = inputs[None, :, None]
  elif is_2d:
    inputs = inputs[:, :, None]
  if method == 'nearest':
    outputs = resize_nearest_v1(inputs, n_timesteps, not add_endpoint)
  elif method == 'linear':
    outputs = resize_bilinear_v1(inputs, n_timesteps, not add_endpoint,
                                 tf_index_math)
  elif method == 'cubic':
    outputs = resize_bicubic_v1(inputs, n_timesteps, not add_endpoint)
  elif method == 'window':
    outputs = upsample_with_windows(inputs, n_timesteps, add_endpoint, dtype)
  else:
    raise ValueError('Method ({}) is invalid. Must be one of {}.'.format(
        method, "['nearest', 'linear', 'cubic', 'window']"))
  if is_1d:
    outputs = outputs[0, :, 0]
  elif is_2d:
    outputs = outputs[:, :, 0]
  elif is_4d:
    outputs = outputs.reshape(shape_4d[0], n_timesteps, shape_4d[2], shape_4d[3])
  return outputs


# ----------------------------------------------------------------------------
# Harmonic synthesis (core.py:797-1111)
# ----------------------------------------------------------------------------
def _cumsum_sequential(x, axis, dtype):
  """np.cumsum accumulates sequentially in the array dtype (as TF-CPU does)."""
  return np.cumsum(x, axis=axis, dtype=dtype)


def angular_cumsum(angular_frequency, chunk_size=1000, dtype=np.float64):
  """core.py:799-866."""
  af = _as(angular_frequency, dtype)
  n_batch, n_time = af.shape[0], af.shape[1]
  rest = af.shape[2:]
  two_pi = np.asarray(TWO_PI, dtype)
  remainder = n_time % chunk_size
  if remainder:
    pad_amount = chunk_size - remainder
    af = np.concatenate(
        [af, np.zeros((n_batch, pad_amount) + rest, dtype)], axis=1)
  length = af.shape[1]
  n_chunks = length // chunk_size
  chunks = af.reshape((n_batch, n_chunks, chunk_size) + rest)
  phase = _cumsum_sequential(chunks, 2, dtype)
  offsets = np.mod(phase[:, :, -1:, ...], two_pi).astype(dtype)
  offsets = np.concatenate(
      [np.zeros_like(offsets[:, :1]), offsets], axis=1)[:, :-1]
  offsets = np.mod(_cumsum_sequential(offsets, 1, dtype), two_pi).astype(dtype)
  phase = (phase + offsets).astype(dtype)
  phase = np.mod(phase, two_pi).astype(dtype)
  phase = phase.reshape((n_batch, length) + rest)
  if remainder:
    phase = phase[:, :n_time]
  return phase


def get_harmonic_frequencies(frequencies, n_harmonics, dtype=np.float64):
  """core.py:1028-1045: f0 * [1..K]."""
  frequencies = _as(frequencies, dtype)
  f_ratios = np.linspace(1.0, float(n_harmonics), int(n_harmonics)).astype(dtype)
  return (frequencies * f_ratios[None, None, :]).astype(dtype)


def remove_above_nyquist(frequency_envelopes, amplitude_envelopes,
                         sample_rate=16000):
  """core.py:869-891."""
  return np.where(frequency_envelopes >= sample_rate / 2.0,
                  np.zeros_like(amplitude_envelopes), amplitude_envelopes)


def normalize_harmonics(harmonic_distribution, f0_hz=None, sample_rate=None,
                        dtype=np.float64):
  """core.py:894-907."""
  hd = _as(harmonic_distribution, dtype)
  if sample_rate is not None and f0_hz is not None:
    n_harmonics = int(hd.shape[-1])
    hf = get_harmonic_frequencies(f0_hz, n_harmonics, dtype)
    hd = remove_above_nyquist(hf, hd, sample_rate)
  total = np.sum(hd, axis=-1, keepdims=True, dtype=dtype)
  return safe_divide(hd, total).astype(dtype)


def oscillator_bank(frequency_envelopes, amplitude_envelopes, sample_rate=16000,
                    sum_sinusoids=True, use_angular_cumsum=False,
                    dtype=np.float64, nyquist_mask=None):
  """core.py:911-962.

  nyquist_mask (bool, True = silenced): when given it replaces the comparison
  `frequency_envelopes >= sample_rate / 2` - used by the float64 arbiter to take
  that yes/no DECISION exactly as the reference's float32 arithmetic would, so
  the arbiter and a float32 implementation cannot disagree by a whole
  oscillator on a sample where f_k(t) is within an ulp of Nyquist.
  """
  fe = _as(frequency_envelopes, dtype)
  ae = _as(amplitude_envelopes, dtype)
  if nyquist_mask is None:
    ae = remove_above_nyquist(fe, ae, sample_rate)
  else:
    ae = np.where(nyquist_mask, np.zeros_like(ae), ae)
  omegas = (fe * np.asarray(TWO_PI, dtype)).astype(dtype)
  omegas = (omegas / np.asarray(float(sample_rate), dtype)).astype(dtype)
  if use_angular_cumsum:
    phases = angular_cumsum(omegas, dtype=dtype)
  else:
    phases = _cumsum_sequential(omegas, 1, dtype)
  wavs = np.sin(phases).astype(dtype)
  audio = (ae * wavs).astype(dtype)
  if sum_sinusoids:
    audio = np.sum(audio, axis=-1, dtype=dtype)
  return audio


def harmonic_synthesis(frequencies, amplitudes, harmonic_shifts=None,
                       harmonic_distribution=None, n_samples=64000,
                       sample_rate=16000, amp_resample_method='window',
                       use_angular_cumsum=False, dtype=np.float64,
                       tf_index_math=False, mask_in_float32=True):
  """core.py:1048-1111.

  mask_in_float32: in float64 mode, evaluate the audio-rate Nyquist decision
  (core.py:888-890) on the float32 frequency envelopes the reference would have
  (f0 * k and the bilinear lerp both rounded to float32, no FMA contraction).
  """
  frequencies = _as(frequencies, dtype)
  amplitudes = _as(amplitudes, dtype)
  if harmonic_distribution is not None:
    harmonic_distribution = _as(harmonic_distribution, dtype)
    n_harmonics = int(harmonic_distribution.shape[-1])
  elif harmonic_shifts is not None:
    harmonic_shifts = _as(harmonic_shifts, dtype)
    n_harmonics = int(harmonic_shifts.shape[-1])
  else:
    n_harmonics = 1
  hf = get_harmonic_frequencies(frequencies, n_harmonics, dtype)
  if harmonic_shifts is not None:
    hf = (hf * (1.0 + harmonic_shifts)).astype(dtype)
  if harmonic_distribution is not None:
    ha = (amplitudes * harmonic_distribution).astype(dtype)
  else:
    ha = amplitudes
  fe = resample(hf, n_samples, dtype=dtype, tf_index_math=tf_index_math)
  ae = resample(ha, n_samples, method=amp_resample_method, dtype=dtype,
                tf_index_math=tf_index_math)
  mask = None
  if mask_in_float32 and dtype != np.float32:
    hf32 = get_harmonic_frequencies(frequencies.astype(np.float32), n_harmonics,
                                    np.float32)
    if harmonic_shifts is not None:
      hf32 = (hf32 * (np.float32(1.0) + harmonic_shifts.astype(np.float32))
              ).astype(np.float32)
    fe32 = resample(hf32, n_samples, dtype=np.float32,
                    tf_index_math=tf_index_math)
    mask = fe32 >= np.float32(sample_rate / 2.0)
  return oscillator_bank(fe, ae, sample_rate=sample_rate,
                         use_angular_cumsum=use_angular_cumsum, dtype=dtype,
                         nyquist_mask=mask)


def harmonic_oscillator_bank(frequency, amplitude_envelopes, initial_phase=None,
                             sample_rate=16000, use_angular_cumsum=True,
                             dtype=np.float64):
  """core.py:966-1025."""
  frequency = _as(frequency, dtype)
  ae = _as(amplitude_envelopes, dtype)
  omega = (frequency * np.asarray(TWO_PI, dtype)).astype(dtype)
  omega = (omega / np.asarray(float(sample_rate), dtype)).astype(dtype)
  if use_angular_cumsum:
    phases = angular_cumsum(omega, dtype=dtype)
  else:
    phases = _cumsum_sequential(omega, 1, dtype)
  if initial_phase is None:
    initial_phase = np.zeros([phases.shape[0], 1, 1], dtype)
  phases = (phases + _as(initial_phase, dtype)).astype(dtype)
  final_phase = phases[:, -1:, 0:1]
  n_harmonics = int(ae.shape[-1])
  f_ratios = np.linspace(1.0, float(n_harmonics), n_harmonics).astype(dtype)
  phases = (phases * f_ratios[None, None, :]).astype(dtype)
  audio = np.sum((ae * np.sin(phases)).astype(dtype), axis=-1, dtype=dtype)
  return audio, final_phase


def streaming_harmonic_synthesis(frequencies, amplitudes,
                                 harmonic_distribution=None, initial_phase=None,
                                 n_samples=64000, sample_rate=16000,
                                 amp_resample_method='linear', dtype=np.float64):
  """core.py:1114-1164."""
  frequencies = _as(frequencies, dtype)
  amplitudes = _as(amplitudes, dtype)
  if harmonic_distribution is not None:
    hd = normalize_harmonics(_as(harmonic_distribution, dtype), frequencies,
                             sample_rate, dtype)
    ha = (amplitudes * hd).astype(dtype)
  else:
    ha = amplitudes
  fe = resample(frequencies, n_samples, dtype=dtype)
  ae = resample(ha, n_samples, method=amp_resample_method, dtype=dtype)
  return harmonic_oscillator_bank(fe, ae, initial_phase, sample_rate, dtype=dtype)


# ----------------------------------------------------------------------------
# Time-varying FIR / filtered noise (core.py:1316-1655)
# ----------------------------------------------------------------------------
def get_fft_size(frame_size, ir_size, power_of_2=True):
  """core.py:1317-1335."""
  convolved_frame_size = ir_size + frame_size - 1
  if not power_of_2:
    raise NotImplementedError
  return int(2**np.ceil(np.log2(convolved_frame_size)))


def crop_and_compensate_delay(audio, audio_size, ir_size, padding,
                              delay_compensation):
  """core.py:1338-1379."""
  if padding == 'valid':
    crop_size = ir_size + audio_size - 1
  elif padding == 'same':
    crop_size = audio_size
  else:
    raise ValueError('Padding must be \'valid\' or \'same\', instead '
                     'of {}.'.format(padding))
  total_size = int(audio.shape[-1])
  crop = total_size - crop_size
  start = ((ir_size - 1) // 2 - 1 if delay_compensation < 0
           else delay_compensation)
  end = crop - start
  return audio[:, start:-end]


def fft_convolve(audio, impulse_response, padding='same', delay_compensation=-1,
                 dtype=np.float64):
  """core.py:1382-1473 - literal framed FFT convolution + overlap-add."""
  audio = _as(audio, dtype)
  ir = _as(impulse_response, dtype)
  batch_size, audio_size = audio.shape
  if ir.ndim == 2:
    ir = ir[:, None, :]
  if ir.shape[0] == 1 and batch_size > 1:
    ir = np.tile(ir, [batch_size, 1, 1])
  batch_size_ir, n_ir_frames, ir_size = ir.shape
  if batch_size != batch_size_ir:
    raise ValueError('Batch size of audio ({}) and impulse response ({}) must '
                     'be the same.'.format(batch_size, batch_size_ir))
  frame_size = int(np.ceil(audio_size / n_ir_frames))
  hop_size = frame_size
  audio_frames = frame_pad_end(audio, frame_size, hop_size)
  n_audio_frames = int(audio_frames.shape[1])
  if n_audio_frames != n_ir_frames:
    raise ValueError(
        'Number of Audio frames ({}) and impulse response frames ({}) do not '
        'match. For small hop size = ceil(audio_size / n_ir_frames), '
        'number of impulse response frames must be a multiple of the audio '
        'size.'.format(n_audio_frames, n_ir_frames))
  fft_size = get_fft_size(frame_size, ir_size, power_of_2=True)
  cdtype = np.complex64 if dtype == np.float32 else np.complex128
  audio_fft = np.fft.rfft(audio_frames, fft_size).astype(cdtype)
  ir_fft = np.fft.rfft(ir, fft_size).astype(cdtype)
  audio_ir_fft = (audio_fft * ir_fft).astype(cdtype)
  audio_frames_out = np.fft.irfft(audio_ir_fft).astype(dtype)
  audio_out = overlap_and_add(audio_frames_out, hop_size)
  return crop_and_compensate_delay(audio_out, audio_size, ir_size, padding,
                                   delay_compensation)


def apply_window_to_impulse_response(impulse_response, window_size=0,
                                     causal=False, dtype=np.float64):
  """core.py:1477-1531."""
  ir = _as(impulse_response, dtype)
  if causal:
    ir = np.fft.fftshift(ir, axes=-1)
  ir_size = int(ir.shape[-1])
  if (window_size <= 0) or (window_size > ir_size):
    window_size = ir_size
  window = hann_window(window_size, dtype)
  padding = ir_size - window_size
  if padding > 0:
    half_idx = (window_size + 1) // 2
    window = np.concatenate(
        [window[half_idx:], np.zeros([padding], dtype), window[:half_idx]])
  else:
    window = np.fft.fftshift(window, axes=-1)
  ir = (window * ir).astype(dtype)
  if padding > 0:
    first_half_start = (ir_size - (half_idx - 1)) + 1
    second_half_end = half_idx + 1
    ir = np.concatenate(
        [ir[..., first_half_start:], ir[..., :second_half_end]], axis=-1)
  else:
    ir = np.fft.fftshift(ir, axes=-1)
  return ir


def frequency_impulse_response(magnitudes, window_size=0, dtype=np.float64):
  """core.py:1534-1565: irfft of real magnitudes, then window + causal form."""
  magnitudes = _as(magnitudes, dtype)
  ir = np.fft.irfft(magnitudes.astype(
      np.complex64 if dtype == np.float32 else np.complex128)).astype(dtype)
  return apply_window_to_impulse_response(ir, window_size, dtype=dtype)


def frequency_filter(audio, magnitudes, window_size=0, padding='same',
                     dtype=np.float64):
  """core.py:1628-1655."""
  ir = frequency_impulse_response(magnitudes, window_size, dtype)
  return fft_convolve(audio, ir, padding=padding, dtype=dtype)


# ----------------------------------------------------------------------------
# Processors (synths.py:55-196, processors.py:162-176), functional form
# ----------------------------------------------------------------------------
def harmonic_get_controls(amplitudes, harmonic_distribution, f0_hz,
                          sample_rate=16000, scale=True,
                          normalize_below_nyquist=True, dtype=np.float64):
  """synths.py:94-121 with scale_fn=exp_sigmoid (scale=True) or None."""
  amplitudes = _as(amplitudes, dtype)
  hd = _as(harmonic_distribution, dtype)
  f0_hz = _as(f0_hz, dtype)
  if scale:
    amplitudes = exp_sigmoid(amplitudes, dtype=dtype)
    hd = exp_sigmoid(hd, dtype=dtype)
  hd = normalize_harmonics(
      hd, f0_hz, sample_rate if normalize_below_nyquist else None, dtype)
  return {'amplitudes': amplitudes, 'harmonic_distribution': hd,
          'f0_hz': f0_hz}


def harmonic_get_signal(amplitudes, harmonic_distribution, f0_hz,
                        n_samples=64000, sample_rate=16000,
                        amp_resample_method='window', use_angular_cumsum=False,
                        dtype=np.float64, tf_index_math=False):
  """synths.py:123-146."""
  return harmonic_synthesis(
      frequencies=f0_hz, amplitudes=amplitudes,
      harmonic_distribution=harmonic_distribution, n_samples=n_samples,
      sample_rate=sample_rate, amp_resample_method=amp_resample_method,
      use_angular_cumsum=use_angular_cumsum, dtype=dtype,
      tf_index_math=tf_index_math)


def noise_get_controls(magnitudes, initial_bias=-5.0, scale=True,
                       dtype=np.float64):
  """synths.py:165-179."""
  magnitudes = _as(magnitudes, dtype)
  if scale:
    magnitudes = exp_sigmoid(
        (magnitudes + np.asarray(initial_bias, dtype)).astype(dtype),
        dtype=dtype)
  return {'magnitudes': magnitudes}


def noise_get_signal(magnitudes, noise, window_size=257, dtype=np.float64):
  """synths.py:181-196 with the uniform noise injected by the caller."""
  return frequency_filter(noise, magnitudes, window_size=window_size,
                          dtype=dtype)


def add_get_signal(signal_one, signal_two):
  """processors.py:174-176."""
  return signal_one + signal_two


def decoder(amps, harmonic_distribution, f0_hz, noise_magnitudes, noise,
            n_samples=64000, sample_rate=16000, window_size=0,
            dtype=np.float64):
  """The `ae.gin` DAG (training/gin/models/ae.gin:47-72):
  Harmonic -> FilteredNoise -> Add, from raw network outputs."""
  hc = harmonic_get_controls(amps, harmonic_distribution, f0_hz, sample_rate,
                             dtype=dtype)
  harm = harmonic_get_signal(n_samples=n_samples, sample_rate=sample_rate,
                             dtype=dtype, **hc)
  nc = noise_get_controls(noise_magnitudes, dtype=dtype)
  nz = noise_get_signal(nc['magnitudes'], noise, window_size, dtype=dtype)
  return {'harmonic': {'signal': harm, 'controls': hc},
          'filtered_noise': {'signal': nz, 'controls': nc},
          'add': {'signal': add_get_signal(nz, harm)}}


# ----------------------------------------------------------------------------
# Multi-scale spectral loss (losses.py:102-243, spectral_ops.py:34-70)
# ----------------------------------------------------------------------------
def stft_mag(audio, frame_size, overlap=0.75):
  """|tf.signal.stft(frame_length, frame_step, pad_end=True)| with the default
  periodic Hann window (spectral_ops.py:34-47, 67-70)."""
  audio = np.asarray(audio, np.float64)
  step = int(frame_size * (1.0 - overlap))
  frames = frame_pad_end(audio, frame_size, step)
  window = hann_window(frame_size, np.float64)
  return np.abs(np.fft.rfft(frames * window, frame_size, axis=-1))


def spectral_loss(target_audio, audio, fft_sizes=(2048, 1024, 512, 256, 128, 64),
                  mag_weight=1.0, logmag_weight=0.0):
  """SpectralLoss.call with loss_type='L1' (losses.py:194-243)."""
  loss = 0.0
  for size in fft_sizes:
    t = stft_mag(target_audio, size)
    v = stft_mag(audio, size)
    if mag_weight > 0:
      loss += mag_weight * np.mean(np.abs(t - v))
    if logmag_weight > 0:
      loss += logmag_weight * np.mean(np.abs(safe_log(t) - safe_log(v)))
  return loss


# ----------------------------------------------------------------------------
# Philox4x32-10 restatement (Salmon et al. 2011, Random123), used to check the
# in-kernel noise generator bit-for-bit.  Not part of the reference (the
# reference draws tf.random.uniform, synths.py:192-193, unseeded).
# ----------------------------------------------------------------------------
_PHILOX_M0 = np.uint64(0xD2511F53)
_PHILOX_M1 = np.uint64(0xCD9E8D57)
_PHILOX_W0 = np.uint32(0x9E3779B9)
_PHILOX_W1 = np.uint32(0xBB67AE85)


def philox4x32_10(counter, key):
  """counter: uint32 [..., 4]; key: uint32 [..., 2] -> uint32 [..., 4]."""
  c = [counter[..., i].astype(np.uint32) for i in range(4)]
  k0 = np.asarray(key[..., 0], np.uint32)
  k1 = np.asarray(key[..., 1], np.uint32)
  with np.errstate(over='ignore'):
    for _ in range(10):
      p0 = c[0].astype(np.uint64) * _PHILOX_M0
      p1 = c[2].astype(np.uint64) * _PHILOX_M1
      hi0 = (p0 >> np.uint64(32)).astype(np.uint32)
      lo0 = p0.astype(np.uint32)
      hi1 = (p1 >> np.uint64(32)).astype(np.uint32)
      lo1 = p1.astype(np.uint32)
      c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
      k0 = (k0 + _PHILOX_W0).astype(np.uint32)
      k1 = (k1 + _PHILOX_W1).astype(np.uint32)
  return np.stack(c, axis=-1)


def philox_uniform_noise(batch_size, n_samples, seed, offset=0):
  """The in-kernel noise stream of ddsp_b200 (see csrc/philox.cuh):
  counter = (sample_index // 4, batch_index, offset_lo, offset_hi),
  key = (seed_lo, seed_hi); 23 mantissa bits -> [1,2) -> 2x-3 in [-1, 1)."""
  n4 = -(-n_samples // 4)
  ctr = np.zeros((batch_size, n4, 4), np.uint32)
  ctr[..., 0] = np.arange(n4, dtype=np.uint32)[None, :]
  ctr[..., 1] = np.arange(batch_size, dtype=np.uint32)[:, None]
  ctr[..., 2] = np.uint32(offset & 0xFFFFFFFF)
  ctr[..., 3] = np.uint32((offset >> 32) & 0xFFFFFFFF)
  key = np.zeros((batch_size, n4, 2), np.uint32)
  key[..., 0] = np.uint32(seed & 0xFFFFFFFF)
  key[..., 1] = np.uint32((seed >> 32) & 0xFFFFFFFF)
  r = philox4x32_10(ctr, key).reshape(batch_size, n4 * 4)[:, :n_samples]
  bits = (r >> np.uint32(9)) | np.uint32(0x3F800000)
  f = bits.view(np.float32)
  return (np.float32(2.0) * f - np.float32(3.0)).astype(np.float32)


# ----------------------------------------------------------------------------
# Frequency scaling of network outputs and the Sinusoidal synth
# (core.py:219-348, 414-507; synths.py:260-323)
# ----------------------------------------------------------------------------
def logb(x, base=2.0, eps=1e-5, dtype=np.float64):
  """core.logb (core.py:219-221)."""
  return safe_divide(safe_log(np.asarray(x, dtype), eps),
                     safe_log(np.asarray(base, dtype), eps), eps)


def midi_to_hz(notes, dtype=np.float64):
  """core.midi_to_hz (core.py:280-297)."""
  return 440.0 * (2.0 ** ((np.asarray(notes, dtype) - 69.0) / 12.0))


def hz_to_midi(frequencies, dtype=np.float64):
  """core.hz_to_midi (core.py:300-306)."""
  f = np.asarray(frequencies, dtype)
  notes = 12.0 * (logb(f, 2.0, dtype=dtype) - logb(440.0, 2.0, dtype=dtype)) + 69.0
  return np.where(f <= 0.0, 0.0, notes)


def unit_to_hz(unit, hz_min, hz_max, dtype=np.float64):
  """core.unit_to_hz / unit_to_midi (core.py:309-336), clip=False."""
  midi_min, midi_max = hz_to_midi(hz_min, dtype), hz_to_midi(hz_max, dtype)
  return midi_to_hz(midi_min + (midi_max - midi_min) * np.asarray(unit, dtype), dtype)


def frequencies_sigmoid(freqs, depth=1, hz_min=0.0, hz_max=8000.0, dtype=np.float64):
  """core.frequencies_sigmoid (core.py:460-507)."""
  freqs = np.asarray(freqs, dtype)
  if freqs.ndim == 3:
    b, t, c = freqs.shape
    freqs = freqs.reshape(b, t, c // depth, depth)
  else:
    depth = freqs.shape[-1]
  f_probs = 1.0 / (1.0 + np.exp(-freqs))
  hz_scales = []
  hz_min_copy = hz_min
  remainder = hz_max - hz_min
  scale_factor = remainder**(1.0 / depth)
  for i in range(depth):
    if i == depth - 1:
      hz_max = remainder
      hz_min = hz_min_copy
    else:
      hz_max = remainder * (1.0 - 1.0 / scale_factor)
      hz_min = 0
      remainder -= hz_max
    hz_scales.append(unit_to_hz(f_probs[..., i], hz_min, hz_max, dtype))
  return np.sum(np.stack(hz_scales, axis=-1), axis=-1)


def frequencies_softmax(freqs, depth=1, hz_min=20.0, hz_max=8000.0, dtype=np.float64):
  """core.frequencies_softmax (core.py:424-457)."""
  freqs = np.asarray(freqs, dtype)
  if freqs.ndim == 3:
    b, t, c = freqs.shape
    freqs = freqs.reshape(b, t, c // depth, depth)
  else:
    depth = freqs.shape[-1]
  e = np.exp(freqs - freqs.max(axis=-1, keepdims=True))
  f_probs = e / e.sum(axis=-1, keepdims=True)
  unit_bins = np.linspace(0.0, 1.0, depth)[None, None, None, :]
  f_unit = np.sum(unit_bins * f_probs, axis=-1)
  return unit_to_hz(f_unit, hz_min, hz_max, dtype)


def sinusoidal_get_controls(amplitudes, frequencies, sample_rate=16000, depth=1,
                            dtype=np.float64):
  """synths.Sinusoidal.get_controls (synths.py:277-303), default scale fns."""
  amps = exp_sigmoid(np.asarray(amplitudes, dtype))
  freqs = frequencies_sigmoid(frequencies, depth=depth, dtype=dtype)
  amps = remove_above_nyquist(freqs, amps, sample_rate)
  return {'amplitudes': amps, 'frequencies': freqs}


def sinusoidal_get_signal(amplitudes, frequencies, n_samples, sample_rate=16000,
                          amp_resample_method='window', dtype=np.float64):
  """synths.Sinusoidal.get_signal (synths.py:305-323)."""
  amp_env = resample(np.asarray(amplitudes, dtype), n_samples,
                     method=amp_resample_method)
  freq_env = resample(np.asarray(frequencies, dtype), n_samples)
  return oscillator_bank(freq_env, amp_env, sample_rate=sample_rate, dtype=dtype)
