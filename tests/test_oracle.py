"""Pins the CPU oracle (oracle/ddsp_oracle.py) - no GPU needed.

Ports of the reference's REAL tests for the hot path (the ones that compare
values or sizes), plus cross-checks of each TF-op
restatement against an independent implementation (scipy / numpy / closed
forms), plus the known-answer tests the reference lacks for the oscillator.
"""
import numpy as np
import pytest
from scipy import signal

from oracle import ddsp_oracle as o
from oracle import ref_port_torch as rp


# ---- core_test.py:ResampleTest (219-267, 178-198, 295-381) ------------------
def _create_resampled_signals(n_before, n_after, add_endpoint, method):
  before = 1.0 - np.sin(np.linspace(0, np.pi, n_before))
  before = before[np.newaxis, :, np.newaxis]
  after = o.resample(before, n_after, method=method, add_endpoint=add_endpoint,
                     dtype=np.float32, tf_index_math=True)
  return before[0, :, 0], after[0, :, 0]


def _assert_subsampled_close(smaller, larger, add_endpoint, threshold=1e-3):
  n_smaller, n_larger = smaller.size, larger.size
  if add_endpoint:
    n_total = int(n_larger / n_smaller * (n_smaller - 1))
  else:
    n_total = n_larger - 1
  idx = np.linspace(0, n_total, n_smaller).astype(int)
  np.testing.assert_allclose(larger[idx], smaller, atol=threshold)


@pytest.mark.parametrize('add_endpoint', [True, False])
@pytest.mark.parametrize('method', ['linear', 'window'])
def test_upsample_accuracy(add_endpoint, method):
  """core_test.py:242-267 (cubic is outside the hot path)."""
  before, after = _create_resampled_signals(5, 16000, add_endpoint, method)
  _assert_subsampled_close(before, after, add_endpoint)


@pytest.mark.parametrize('add_endpoint', [True, False])
def test_downsample_accuracy_linear(add_endpoint):
  """core_test.py:269-293."""
  before, after = _create_resampled_signals(16000, 5, add_endpoint, 'linear')
  _assert_subsampled_close(after, before, add_endpoint)


@pytest.mark.parametrize('dimensions', [1, 2, 4])
def test_window_only_allows_3d_inputs(dimensions):
  """core_test.py:178-198."""
  with pytest.raises(ValueError):
    o.upsample_with_windows(np.ones([5] * dimensions), 16000)


@pytest.mark.parametrize('add_endpoint', [True, False])
def test_window_checks_for_downsampling(add_endpoint):
  """core_test.py:295-312."""
  with pytest.raises(ValueError):
    o.upsample_with_windows(np.ones([1, 16000, 1]), 5, add_endpoint)


def test_window_disallows_noninteger_upsampling_ratios():
  """core_test.py:314-381."""
  o.upsample_with_windows(np.ones([1, 5, 1]), 15)           # 15 % 5 == 0
  with pytest.raises(ValueError):
    o.upsample_with_windows(np.ones([1, 5, 1]), 16)
  o.upsample_with_windows(np.ones([1, 5, 1]), 16, add_endpoint=False)  # 16 % 4
  with pytest.raises(ValueError):
    o.upsample_with_windows(np.ones([1, 5, 1]), 15, add_endpoint=False)


def test_resample_invalid_method():
  with pytest.raises(ValueError):
    o.resample(np.ones([1, 5, 1]), 10, method='bogus')


def test_window_upsample_is_two_tap_raised_cosine():
  """literal Hann OLA == x[i] w[hop+r] + x[i+1] w[r]."""
  rng = np.random.default_rng(0)
  x = rng.standard_normal((2, 10, 3))
  n, hop = 640, 64
  y = o.upsample_with_windows(x, n)
  t = np.arange(n)
  i, r = t // hop, t % hop
  xe = np.concatenate([x, x[:, -1:]], 1)
  w1 = 0.5 - 0.5 * np.cos(np.pi * r / hop)
  y2 = xe[:, i] * (1 - w1)[None, :, None] + xe[:, i + 1] * w1[None, :, None]
  assert np.abs(y - y2).max() < 1e-14


def test_bilinear_matches_torch_interpolate_semantics():
  """v1 bilinear, align_corners=False, no half-pixel: src = t * in/out."""
  rng = np.random.default_rng(1)
  x = rng.standard_normal((2, 7, 3))
  y = o.resize_bilinear_v1(x, 28)
  for t in range(28):
    src = t * 7 / 28
    lo = int(np.floor(src))
    hi = min(lo + 1, 6)
    want = x[:, lo] + (x[:, hi] - x[:, lo]) * (src - lo)
    np.testing.assert_allclose(y[:, t], want, atol=1e-14)
  # exact and TF float32 index math coincide for power-of-two ratios
  y_tf = o.resize_bilinear_v1(x, 28, tf_index_math=True)
  assert np.abs(y - y_tf).max() < 1e-14


# ---- core_test.py:HarmonicSynthTest -----------------------------------------
@pytest.mark.parametrize('sample_rate', [4000, 16000, 44100])
def test_silent_above_nyquist(sample_rate):
  """core_test.py:484-503."""
  nyquist = sample_rate / 2
  freqs = np.array([1.1, 1.5, 2.0]) * nyquist
  ones = np.ones([2, 16000, 3])
  for dtype in (np.float32, np.float64):
    wav = o.oscillator_bank(ones * freqs, ones, sample_rate=sample_rate,
                            dtype=dtype)
    np.testing.assert_allclose(wav, np.zeros_like(wav))


@pytest.mark.parametrize('sum_sinusoids', [True, False])
def test_oscillator_bank_shape_is_correct(sum_sinusoids):
  """core_test.py:460-482."""
  ones = np.ones([2, 1600, 3])
  wav = o.oscillator_bank(ones * np.array([1.0, 1.5, 2.0]) * 400.0, ones,
                          sum_sinusoids=sum_sinusoids)
  assert list(wav.shape) == ([2, 1600] if sum_sinusoids else [2, 1600, 3])


def test_oscillator_known_answer_constant_frequency():
  """The non-vacuous version of core_test.py:421-458: inclusive cumsum means
  phase(n) = omega (n + 1)."""
  n, sr = 16000, 16000
  f = np.array([440.0, 880.0, 1234.5])
  a = np.array([0.5, 0.3, 0.2])
  ones = np.ones([1, n, 3])
  wav = o.oscillator_bank(ones * f, ones * a, sample_rate=sr)[0]
  t = (np.arange(n) + 1) / sr
  want = (a[None] * np.sin(2 * np.pi * f[None] * t[:, None])).sum(-1)
  assert np.abs(wav - want).max() < 1e-9


def test_harmonic_synthesis_known_answer_and_float32_envelope():
  """harmonic_synthesis at constant f0 (the intent of core_test.py:505-589)
  and the reference-order float32 drift vs the float64 arbiter."""
  n, frames, sr = 16000, 250, 16000
  f0 = np.full((1, frames, 1), 220.0)
  amp = np.ones((1, frames, 1))
  hd = np.full((1, frames, 4), 0.25)
  wav = o.harmonic_synthesis(f0, amp, harmonic_distribution=hd, n_samples=n)[0]
  t = (np.arange(n) + 1) / sr
  want = sum(0.25 * np.sin(2 * np.pi * 220.0 * k * t) for k in range(1, 5))
  assert np.abs(wav - want).max() < 1e-9
  wav32 = o.harmonic_synthesis(f0, amp, harmonic_distribution=hd, n_samples=n,
                               dtype=np.float32)[0]
  drift = np.abs(wav32 - want).max()
  assert 1e-4 < drift < 0.5   # the reference's own float32 phase drift
  wav_ang = o.harmonic_synthesis(f0, amp, harmonic_distribution=hd,
                                 n_samples=n, dtype=np.float32,
                                 use_angular_cumsum=True)[0]
  assert np.abs(wav_ang - want).max() < 1e-1   # chunked float32 still drifts


def test_angular_cumsum_equals_cumsum_mod_2pi():
  rng = np.random.default_rng(3)
  w = rng.uniform(0, 0.5, (2, 2500, 3))
  a = o.angular_cumsum(w, chunk_size=1000)
  b = np.mod(np.cumsum(w, axis=1), 2 * np.pi)
  d = np.abs(a - b)
  assert np.minimum(d, 2 * np.pi - d).max() < 1e-9


def test_normalize_harmonics_nyquist():
  """core_test.py:104-142 (get_harmonic_frequencies / normalize_harmonics)."""
  f0 = np.array([[[1000.0], [3000.0], [4500.0], [9000.0]]])
  hd = np.ones((1, 4, 3))
  out = o.normalize_harmonics(hd, f0, 16000)
  np.testing.assert_allclose(out[0, 0], [1 / 3] * 3)
  np.testing.assert_allclose(out[0, 1], [0.5, 0.5, 0.0])
  np.testing.assert_allclose(out[0, 2], [1.0, 0.0, 0.0])
  np.testing.assert_allclose(out[0, 3], [0.0, 0.0, 0.0])   # safe_divide: 0/1e-7


def test_exp_sigmoid_limits():
  x = np.array([-100.0, 0.0, 100.0])
  y = o.exp_sigmoid(x)
  np.testing.assert_allclose(y, [1e-7, 2.0 * 0.5**np.log(10.0) + 1e-7, 2.0 + 1e-7],
                             rtol=1e-12)


# ---- core_test.py:FiniteImpulseResponseTest ---------------------------------
@pytest.mark.parametrize('audio_size,ir_size', [(1000, 10), (10, 100)])
def test_fft_convolve_is_accurate(audio_size, ir_size):
  """core_test.py:730-757."""
  audio = np.ones([1, audio_size], np.float32)
  ir = np.ones([1, ir_size], np.float32)
  out = o.fft_convolve(audio, ir, padding='valid', delay_compensation=0,
                       dtype=np.float32)[0]
  want = signal.fftconvolve(audio[0], ir[0])
  assert np.abs(want - out).mean() <= 1e-3


def test_fft_convolve_random_vs_scipy():
  rng = np.random.default_rng(5)
  audio = rng.standard_normal((2, 500))
  ir = rng.standard_normal((2, 37))
  out = o.fft_convolve(audio, ir, padding='valid', delay_compensation=0)
  for b in range(2):
    np.testing.assert_allclose(out[b], signal.fftconvolve(audio[b], ir[b]),
                               atol=1e-10)


@pytest.mark.parametrize('gain', [1.0, 0.1])
def test_delay_compensation_corrects_group_delay(gain):
  """core_test.py:759-785."""
  audio = np.random.default_rng(0).standard_normal((1, 1000)).astype(np.float32)
  mags = gain * np.ones([1, 1025], np.float32)
  ir = o.frequency_impulse_response(mags, 257, dtype=np.float32)
  out = o.fft_convolve(audio, ir, padding='same', dtype=np.float32)[0]
  assert np.abs(gain * audio[0] - out).mean() <= 1e-3


def test_fft_convolve_checks():
  """core_test.py:787-823."""
  audio = np.random.default_rng(0).standard_normal((1, 1000))
  with pytest.raises(ValueError):
    o.fft_convolve(audio, np.concatenate([audio, audio], 0))
  for padding in ('same', 'valid'):
    assert o.fft_convolve(audio, audio, padding=padding).shape[0] == 1
  for padding in ('', 'saaammmeee'):
    with pytest.raises(ValueError):
      o.fft_convolve(audio, audio, padding=padding)
  for n_frames in (1010, 999):
    with pytest.raises(ValueError):
      o.fft_convolve(audio, np.zeros((1, n_frames, 1000)))


@pytest.mark.parametrize('fft_size,window_size', [(2048, 0), (2048, 257),
                                                  (1024, 22), (1024, 2048)])
def test_frequency_impulse_response_gives_correct_size(fft_size, window_size):
  """core_test.py:825-855."""
  mags = np.random.default_rng(0).uniform(size=(1, fft_size // 2 + 1))
  ir = o.frequency_impulse_response(mags, window_size)
  target = fft_size
  if target > window_size >= 1:
    target = window_size
    target -= int(target % 2 == 0)
  assert ir.shape[-1] == target


@pytest.mark.parametrize('n_freq,n_frames,window_size', [
    (1025, 0, 0), (1025, 0, 257), (513, 1, 257), (513, 13, 257), (513, 1000, 257)])
def test_frequency_filter_gives_correct_size(n_freq, n_frames, window_size):
  """core_test.py:857-886."""
  rng = np.random.default_rng(0)
  audio = rng.standard_normal((1, 1000))
  shape = (1, n_frames, n_freq) if n_frames > 0 else (1, n_freq)
  out = o.frequency_filter(audio, rng.uniform(size=shape), window_size)
  assert out.shape[-1] == 1000


def test_direct_form_fir_equals_fft_convolve():
  """framed FFT conv + OLA + crop == time-varying FIR whose taps
  are chosen by the INPUT sample's frame."""
  rng = np.random.default_rng(2)
  B, F, nb, N = 2, 10, 65, 640
  mags = rng.uniform(0.1, 1, (B, F, nb))
  noise = rng.uniform(-1, 1, (B, N))
  ref = o.frequency_filter(noise, mags, window_size=0)
  ir = o.frequency_impulse_response(mags, 0)
  S = ir.shape[-1]
  assert S == 128 and np.abs(ir[..., 0]).max() == 0.0
  start, frame = (S - 1) // 2 - 1, N // F
  out = np.zeros((B, N))
  for t in range(N):
    for m in range(S):
      p = t + start - m
      if 0 <= p < N:
        out[:, t] += ir[:, p // frame, m] * noise[:, p]
  assert np.abs(out - ref).max() < 1e-12
  # flat magnitudes g: the 128-tap filter is g * delta delayed by 2 samples
  flat = o.frequency_filter(noise, np.full((B, F, nb), 0.7), window_size=0)
  assert np.abs(flat[:, 2:] - 0.7 * noise[:, :-2]).max() < 1e-12


def test_quarter_wave_cosine_sums_equal_the_impulse_response():
  """The identities noise_ring's producers rest on (csrc/noise_ring.cuh, section B),
  restated in NumPy with the kernel's own table layout and tap placement: with
  m_k the 65 magnitudes, h0 = irfft(m) splits into even-k / odd-k cosine sums
  (h0[n] = E[n] + O[n], h0[64 - n] = E[n] - O[n]) and the even-k half once more
  about n = 16 (E[n] = EE[n] + EO[n], E[32 - n] = EE[n] - EO[n]); every one of
  the 128 windowed taps is written, tap 128 never (core.py:1476-1519)."""
  nb, S, Q, QP, shift = 65, 128, 32, 36, 64
  rng = np.random.default_rng(0)
  m = rng.uniform(0.0, 2.0, size=nb)
  te = np.zeros((33, QP))
  to = np.zeros((32, QP))
  for k in range(33):
    ck = (1.0 / S) if k in (0, 32) else 2.0 / S
    for n in range(Q + 1):
      te[k, n] = ck * np.cos(2.0 * np.pi * ((2 * k * n) % S) / S)
  for k in range(32):                       # odd-k table, per producer warp
    for col in range(33):
      w, c = col >> 3, col & 7
      n = 16 if col == 32 else (4 * w + c if c < 4 else 32 - 4 * w - (c - 4))
      to[k, col] = 2.0 / S * np.cos(2.0 * np.pi * (((2 * k + 1) * n) % S) / S)
  win = 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(S) / S)
  hr = np.full(S + 4, np.nan)
  hr[S:] = 0.0                               # the row's zero pad
  for iw in range(4):
    c0 = 4 * iw
    ee, eo, oa, ob = np.zeros(4), np.zeros(4), np.zeros(4), np.zeros(4)
    ee16 = o16 = 0.0
    for j2 in range(16):
      m0, m1, m2, m3 = m[4 * j2:4 * j2 + 4]
      ee += m0 * te[2 * j2, c0:c0 + 4]
      eo += m2 * te[2 * j2 + 1, c0:c0 + 4]
      oa += m1 * to[2 * j2, 8 * iw:8 * iw + 4] + m3 * to[2 * j2 + 1, 8 * iw:8 * iw + 4]
      ob += m1 * to[2 * j2, 8 * iw + 4:8 * iw + 8] + m3 * to[2 * j2 + 1, 8 * iw + 4:8 * iw + 8]
      ee16 += m0 * te[2 * j2, 16]
      o16 += m1 * to[2 * j2, 32] + m3 * to[2 * j2 + 1, 32]
    ee += m[64] * te[32, c0:c0 + 4]
    ee16 += m[64] * te[32, 16]
    e_a, e_b = ee + eo, ee - eo              # E[n], E[32 - n]
    vpa = win[shift + c0:shift + c0 + 4] * (e_a + oa)
    vma = win[c0:c0 + 4] * (e_a - oa)
    vpb = win[Q + c0:Q + c0 + 4] * (e_b + ob)          # win[64 + (32 - n)] == win[32 + n]
    vmb = win[shift + Q + c0:shift + Q + c0 + 4] * (e_b - ob)

    def up(base, v):
      hr[base:base + 4] = v

    def down(base, v, first=True):
      if first:
        hr[base] = v[0]
      hr[base - 3:base] = v[3:0:-1]

    up(shift + c0, vpa); down(shift - c0, vpa)
    up(c0, vma); down(S - c0, vma, first=(c0 != 0))
    down(shift + Q - c0, vpb); up(Q + c0, vpb)
    down(Q - c0, vmb); up(shift + Q + c0, vmb)
    if iw == 3:
      hr[shift + 16] = hr[shift - 16] = win[shift + 16] * (ee16 + o16)
      hr[16] = hr[S - 16] = win[16] * (ee16 - o16)
  assert not np.isnan(hr[:S]).any()
  assert np.all(hr[S:] == 0.0)
  want = o.frequency_impulse_response(m[None, None, :], window_size=0, dtype=np.float64)[0, 0]
  np.testing.assert_allclose(hr[:S], want, rtol=0, atol=1e-14)


def test_overlap_and_add_and_frame():
  x = np.arange(10.0)[None]
  fr = o.frame_pad_end(x, 4, 4)
  assert fr.shape == (1, 3, 4) and fr[0, 2].tolist() == [8.0, 9.0, 0.0, 0.0]
  y = o.overlap_and_add(np.ones((1, 3, 4)), 2)
  assert y[0].tolist() == [1, 1, 2, 2, 2, 2, 1, 1]


# ---- Philox and the torch baseline port -------------------------------------
def test_philox_known_answer():
  """Random123 kat_vectors: philox4x32-10, counter = key = 0 and all ones."""
  out = o.philox4x32_10(np.zeros((1, 4), np.uint32), np.zeros((1, 2), np.uint32))
  assert [hex(v) for v in out[0]] == ['0x6627e8d5', '0xe169c58d', '0xbc57ac4c',
                                      '0x9b00dbd8']
  ff = np.full((1, 4), 0xFFFFFFFF, np.uint32)
  out = o.philox4x32_10(ff, ff[:, :2])
  assert [hex(v) for v in out[0]] == ['0x408f276d', '0x41c83b0e', '0xa20bc7c6',
                                      '0x6d5451fd']


def test_philox_uniform_range():
  x = o.philox_uniform_noise(2, 4001, seed=7, offset=1)
  assert x.shape == (2, 4001) and x.dtype == np.float32
  assert x.min() >= -1.0 and x.max() < 1.0 and abs(x.mean()) < 0.05
  assert not np.array_equal(x[0], x[1])


def test_torch_port_matches_oracle():
  """The timed CPU baseline computes the same decoder as the oracle (within the
  reference's own float32 phase-drift envelope at this length)."""
  import torch
  from tests.util import synth_inputs
  inp = synth_inputs(2, 125, 100, 65, 8000, seed=5)
  t = {k: torch.from_numpy(v) for k, v in inp.items()}
  got = rp.decoder(t['amps'], t['harmonic_distribution'], t['f0_hz'],
                   t['noise_magnitudes'], 8000, noise=t['noise']).numpy()
  want = o.decoder(inp['amps'], inp['harmonic_distribution'], inp['f0_hz'],
                   inp['noise_magnitudes'], inp['noise'], n_samples=8000,
                   dtype=np.float64)['add']['signal']
  assert np.abs(got - want).max() < 5e-3 * np.abs(want).max()
  # phase-free parts agree tightly
  nz = rp.noise_signal(rp.noise_controls(t['noise_magnitudes']), 8000, 0,
                       t['noise']).numpy()
  nz_want = o.noise_get_signal(
      o.noise_get_controls(inp['noise_magnitudes'])['magnitudes'],
      inp['noise'], 0)
  assert np.abs(nz - nz_want).max() < 1e-5 * max(np.abs(nz_want).max(), 1e-9)
  a, h = rp.harmonic_controls(t['amps'], t['harmonic_distribution'], t['f0_hz'])
  c = o.harmonic_get_controls(inp['amps'], inp['harmonic_distribution'],
                              inp['f0_hz'])
  np.testing.assert_allclose(h.numpy(), c['harmonic_distribution'], rtol=1e-4,
                             atol=1e-9)
