"""Randomised cross-check of the oracle against the UNMODIFIED reference run on the
NumPy shim (CPU, authoring container only: skipped where /root/reference is absent).

The golden fixtures pin the oracle on the path's configurations; this sweeps the
argument space around them - shapes, paddings, delays, odd / even / degenerate filter
windows, every amplitude resampling method, both phase accumulators, sample rates,
optional arguments - in the reference's "wide" mode (its own code evaluated in float64)
against the oracle's float64 mode.  It is how the odd-window Hann discrepancy was found.
"""
import numpy as np
import pytest

from oracle import ddsp_oracle as o
from oracle import ref_on_shim

pytestmark = pytest.mark.skipif(
    not ref_on_shim.available(),
    reason='reference sources (/root/reference) are only in the authoring container')


@pytest.fixture(scope='module')
def ref():
  return ref_on_shim.load(), ref_on_shim.tf()


def _wide(tf, fn):
  tf.set_wide(True)
  try:
    return ref_on_shim.to_numpy(fn())
  finally:
    tf.set_wide(False)


def _close(got, want, tol, what):
  got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
  assert got.shape == want.shape, (what, got.shape, want.shape)
  if got.size:
    assert np.abs(got - want).max() <= tol * max(1.0, np.abs(want).max()), what


def test_fft_convolve_shapes_paddings_delays(ref):
  ddsp, tf = ref
  rng = np.random.default_rng(123)
  checked = 0
  for _ in range(24):
    b, f = int(rng.integers(1, 3)), int(rng.choice([1, 2, 5, 10, 25]))
    frame, s = int(rng.choice([1, 3, 16, 48, 64])), int(rng.choice([1, 2, 3, 10, 31, 64, 65, 128, 200]))
    pad, dc = str(rng.choice(['same', 'valid'])), int(rng.choice([-1, 0, 1, 5]))
    a = rng.standard_normal((b, f * frame)).astype(np.float32)
    ir = rng.standard_normal((b, f, s)).astype(np.float32)
    try:
      want = _wide(tf, lambda: ddsp.core.fft_convolve(a, ir, padding=pad, delay_compensation=dc))
    except Exception:  # pylint: disable=broad-except
      with pytest.raises(Exception):
        o.fft_convolve(a.astype(np.float64), ir.astype(np.float64), padding=pad,
                       delay_compensation=dc)
      continue
    got = o.fft_convolve(a.astype(np.float64), ir.astype(np.float64), padding=pad,
                         delay_compensation=dc)
    _close(got, want, 1e-9, ('fft_convolve', b, f, frame, s, pad, dc))
    checked += 1
  assert checked >= 12


def test_frequency_filter_windows(ref):
  ddsp, tf = ref
  rng = np.random.default_rng(124)
  for _ in range(20):
    f, frame = int(rng.choice([1, 4, 10])), int(rng.choice([8, 32, 64]))
    nb = int(rng.choice([2, 3, 9, 16, 33, 65, 100, 129, 130, 257]))
    ws = int(rng.choice([0, 1, 2, 3, 7, 8, 50, 51, 64, 65, 257]))
    a = rng.uniform(-1, 1, (1, f * frame)).astype(np.float32)
    m = rng.uniform(0, 1, (1, f, nb)).astype(np.float32)
    want = _wide(tf, lambda: ddsp.core.frequency_filter(a, m, window_size=ws))
    got = o.frequency_filter(a.astype(np.float64), m.astype(np.float64), window_size=ws)
    _close(got, want, 1e-9, ('frequency_filter', f, frame, nb, ws))


def test_harmonic_synthesis_argument_space(ref):
  ddsp, tf = ref
  rng = np.random.default_rng(321)
  for _ in range(14):
    b, f = int(rng.integers(1, 3)), int(rng.choice([2, 5, 10, 25]))
    hop, k = int(rng.choice([4, 16, 64, 100])), int(rng.choice([1, 3, 20, 60]))
    method = str(rng.choice(['window', 'linear', 'nearest', 'cubic']))
    uac, sr = bool(rng.integers(0, 2)), int(rng.choice([16000, 8000, 44100]))
    f0 = rng.uniform(20, sr * 0.45, (b, f, 1)).astype(np.float32)
    amp = rng.uniform(0, 1, (b, f, 1)).astype(np.float32)
    hd = rng.uniform(0, 1, (b, f, k)).astype(np.float32) if rng.integers(0, 4) else None
    shifts = (rng.uniform(-0.05, 0.05, (b, f, k)).astype(np.float32)
              if hd is not None and rng.integers(0, 2) else None)
    # tensors, so that the wide mode widens every operand (a raw float32 array in
    # `1.0 + harmonic_shifts` would be rounded by NumPy before the shim sees it)
    t = lambda x: None if x is None else tf.convert_to_tensor(x)  # noqa: E731
    want = _wide(tf, lambda: ddsp.core.harmonic_synthesis(
        t(f0), t(amp), harmonic_shifts=t(shifts), harmonic_distribution=t(hd),
        n_samples=f * hop, sample_rate=sr, amp_resample_method=method,
        use_angular_cumsum=uac))
    w = lambda x: None if x is None else x.astype(np.float64)  # noqa: E731
    got = o.harmonic_synthesis(w(f0), w(amp), harmonic_shifts=w(shifts),
                               harmonic_distribution=w(hd), n_samples=f * hop,
                               sample_rate=sr, amp_resample_method=method,
                               use_angular_cumsum=uac, dtype=np.float64)
    _close(got, want, 2e-7, ('harmonic_synthesis', b, f, hop, k, method, uac, sr))


def test_controls_oscillators_streaming_and_scalers(ref):
  ddsp, tf = ref
  rng = np.random.default_rng(999)
  t = tf.convert_to_tensor
  for _ in range(8):                                   # Harmonic.get_controls variants
    f, k = int(rng.choice([3, 10])), int(rng.choice([1, 7, 40]))
    scale, nyq, sr = bool(rng.integers(0, 2)), bool(rng.integers(0, 2)), int(rng.choice([16000, 4000]))
    a = rng.standard_normal((1, f, 1)).astype(np.float32)
    h = rng.standard_normal((1, f, k)).astype(np.float32)
    f0 = rng.uniform(0, sr / 2, (1, f, 1)).astype(np.float32)
    if not scale:
      a, h = np.abs(a), np.abs(h)
    syn = ddsp.synths.Harmonic(n_samples=f * 8, sample_rate=sr,
                               scale_fn=ddsp.core.exp_sigmoid if scale else None,
                               normalize_below_nyquist=nyq)
    want = _wide(tf, lambda: syn.get_controls(a, h, f0))
    got = o.harmonic_get_controls(a.astype(np.float64), h.astype(np.float64),
                                  f0.astype(np.float64), sample_rate=sr, scale=scale,
                                  normalize_below_nyquist=nyq, dtype=np.float64)
    for key in ('amplitudes', 'harmonic_distribution', 'f0_hz'):
      _close(got[key], want[key], 1e-12, ('get_controls', key, f, k, scale, nyq, sr))
  for _ in range(8):                                   # oscillator_bank
    b, n, k = int(rng.integers(1, 3)), int(rng.choice([50, 1000, 2500])), int(rng.choice([1, 4, 17]))
    sr, ss, uac = int(rng.choice([16000, 8000])), bool(rng.integers(0, 2)), bool(rng.integers(0, 2))
    fe = rng.uniform(0, sr * 0.6, (b, n, k)).astype(np.float32)
    ae = rng.uniform(0, 1, (b, n, k)).astype(np.float32)
    want = _wide(tf, lambda: ddsp.core.oscillator_bank(t(fe), t(ae), sample_rate=sr,
                                                       sum_sinusoids=ss, use_angular_cumsum=uac))
    got = o.oscillator_bank(fe.astype(np.float64), ae.astype(np.float64), sample_rate=sr,
                            sum_sinusoids=ss, use_angular_cumsum=uac, dtype=np.float64)
    _close(got, want, 1e-8, ('oscillator_bank', b, n, k, sr, ss, uac))
  for _ in range(8):                                   # streaming synthesis, carried phase
    b, f, hop = int(rng.integers(1, 3)), int(rng.choice([1, 4, 10])), int(rng.choice([16, 64]))
    k = int(rng.choice([1, 5, 30]))
    f0 = rng.uniform(50, 2000, (b, f, 1)).astype(np.float32)
    amp = rng.uniform(0, 1, (b, f, 1)).astype(np.float32)
    hd = rng.uniform(0, 1, (b, f, k)).astype(np.float32) if rng.integers(0, 3) else None
    ph = rng.uniform(0, 6.28, (b, 1, 1)).astype(np.float32) if rng.integers(0, 2) else None
    method = str(rng.choice(['linear', 'window']))
    tt = lambda x: None if x is None else t(x)  # noqa: E731
    want = _wide(tf, lambda: list(ddsp.core.streaming_harmonic_synthesis(
        t(f0), t(amp), tt(hd), tt(ph), n_samples=f * hop, sample_rate=16000,
        amp_resample_method=method)))
    w = lambda x: None if x is None else x.astype(np.float64)  # noqa: E731
    got = o.streaming_harmonic_synthesis(w(f0), w(amp), w(hd), w(ph), n_samples=f * hop,
                                         sample_rate=16000, amp_resample_method=method,
                                         dtype=np.float64)
    _close(got[0], want[0], 1e-8, ('streaming audio', b, f, hop, k, method))
    d = np.angle(np.exp(1j * (np.asarray(got[1], np.float64) - np.asarray(want[1], np.float64))))
    assert np.abs(d).max() <= 1e-8
  for _ in range(5):                                   # scaling functions
    x = (4 * rng.standard_normal((2, 5, 7))).astype(np.float32)
    ex, mv, th = float(rng.choice([10.0, 2.0, 5.0])), float(rng.choice([2.0, 1.0])), float(rng.choice([1e-7, 1e-3]))
    _close(o.exp_sigmoid(x.astype(np.float64), ex, mv, th, dtype=np.float64),
           _wide(tf, lambda: ddsp.core.exp_sigmoid(t(x), ex, mv, th)), 1e-12, 'exp_sigmoid')
    depth = int(rng.choice([1, 8, 64]))
    fr = rng.standard_normal((2, 5, 3 * depth)).astype(np.float32)
    _close(o.frequencies_sigmoid(fr.astype(np.float64), depth=depth, dtype=np.float64),
           _wide(tf, lambda: ddsp.core.frequencies_sigmoid(t(fr), depth=depth)), 1e-9, 'sigmoid')
    _close(o.frequencies_softmax(fr.astype(np.float64), depth=depth, dtype=np.float64),
           _wide(tf, lambda: ddsp.core.frequencies_softmax(t(fr), depth=depth)), 1e-9, 'softmax')
  for _ in range(4):                                   # Sinusoidal
    f, k, hop = int(rng.choice([5, 10])), int(rng.choice([1, 4, 9])), int(rng.choice([16, 64]))
    a = rng.standard_normal((1, f, k)).astype(np.float32)
    fr = rng.standard_normal((1, f, k)).astype(np.float32)
    method = str(rng.choice(['window', 'linear']))
    syn = ddsp.synths.Sinusoidal(n_samples=f * hop, sample_rate=16000, amp_resample_method=method)
    want = _wide(tf, lambda: syn(t(a), t(fr)))
    c = o.sinusoidal_get_controls(a.astype(np.float64), fr.astype(np.float64), dtype=np.float64)
    got = o.sinusoidal_get_signal(c['amplitudes'], c['frequencies'], f * hop,
                                  amp_resample_method=method, dtype=np.float64)
    _close(got, want, 1e-8, ('sinusoidal', f, k, hop, method))
