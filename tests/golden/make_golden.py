"""Writes tests/golden/*.npz: outputs of the UNMODIFIED REFERENCE (magenta/ddsp,
/root/reference/ddsp) on seeded inputs, for the decoder path.

How: the reference package is imported as it lies and run on the NumPy stand-in
for its TensorFlow primitives (oracle/tf_shim via oracle/ref_on_shim.py) - twice:
"narrow" (float32, the reference's own arithmetic incl. its sequential float32
phase cumsum) and "wide" (the same reference code with every float32 widened to
float64: the exact value of the reference's formulae, which is what the 1e-4
parity gate is measured against; BASELINE.md section 5 gives the distance between
the two - the reference's own phase-accumulation error).

Needs /root/reference, so it runs in the authoring container only:

  python tests/golden/make_golden.py          # rewrite every fixture
  python tests/golden/make_golden.py --check  # regenerate in memory and compare

The fixtures travel to the GPU box; tests/test_reference_pin.py (CPU: oracle vs
fixtures, and fixtures vs a fresh reference run when the reference is present)
and tests/test_gpu_golden.py (CUDA path vs fixtures) read them.  Inputs are
regenerated from their seeds by tests/util.synth_inputs; an input checksum in
every fixture guards against generator drift.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ref_on_shim               # noqa: E402
from tests.util import synth_inputs          # noqa: E402


def checksum(inp):
  return np.float64(sum(float(np.asarray(v, np.float64).sum()) * (i + 1)
                        for i, v in enumerate(inp[k] for k in sorted(inp))))


def _both(fn):
  """fn() under the narrow and the wide shim -> (narrow, wide) numpy results."""
  tf = ref_on_shim.tf()
  out = []
  for wide in (False, True):
    tf.set_wide(wide)
    try:
      out.append(ref_on_shim.to_numpy(fn()))
    finally:
      tf.set_wide(False)
  return out


def _nyquist_margin(ddsp, inp, n_samples, k, sample_rate=16000):
  """Smallest |f_k(t) - sr/2| of the float32 frequency envelopes: the fixture is
  only meaningful if no oscillator sits within an ulp of the Nyquist decision
  (then narrow and wide would disagree by a whole oscillator at that sample)."""
  hf = ddsp.core.get_harmonic_frequencies(inp['f0_hz'], k)
  fe = ddsp.core.resample(hf, n_samples).numpy().astype(np.float64)
  return float(np.abs(fe - sample_rate / 2.0).min())


def c1_harmonic():
  """BASELINE.json configs[0]: Harmonic only, B=1, 16000 samples, 64 harmonics,
  250 frames (synths.Harmonic defaults: window resampling, Nyquist normalise)."""
  ddsp = ref_on_shim.load()
  seed = 101
  inp = synth_inputs(1, 250, 64, 65, 16000, seed=seed)
  args = (inp['amps'], inp['harmonic_distribution'], inp['f0_hz'])
  assert _nyquist_margin(ddsp, inp, 16000, 64) > 1e-2

  def run(angular):
    h = ddsp.synths.Harmonic(n_samples=16000, use_angular_cumsum=angular)
    return h(*args, return_outputs_dict=True)

  n0, w0 = _both(lambda: run(False))
  n1, _ = _both(lambda: run(True))
  return dict(
      seed=seed, input_checksum=checksum(inp),
      amplitudes=n0['controls']['amplitudes'],
      harmonic_distribution=n0['controls']['harmonic_distribution'],
      audio_ref_f32_cumsum=n0['signal'], audio_ref_f32_angular=n1['signal'],
      audio_ref_wide=w0['signal'].astype(np.float64))


def decoder_small():
  """The ae.gin DAG (ae.gin:47-72) through the reference's ProcessorGroup from raw
  network outputs, B=2, F=25, N=1600, K=100, 65 bands, injected noise."""
  ddsp = ref_on_shim.load()
  tf = ref_on_shim.tf()
  seed = 202
  B, F, K, nb, N = 2, 25, 100, 65, 1600
  inp = synth_inputs(B, F, K, nb, N, seed=seed)
  assert _nyquist_margin(ddsp, inp, N, K) > 1e-2

  def run():
    harm = ddsp.synths.Harmonic(n_samples=N, sample_rate=16000)
    noise = ddsp.synths.FilteredNoise(n_samples=N, window_size=0)
    group = ddsp.processors.ProcessorGroup(dag=[
        (harm, ['amps', 'harmonic_distribution', 'f0_hz']),
        (noise, ['noise_magnitudes']),
        (ddsp.processors.Add(), ['filtered_noise/signal', 'harmonic/signal'])])
    tf.random.inject_uniform(inp['noise'])
    feats = {k: inp[k] for k in ('amps', 'harmonic_distribution', 'f0_hz',
                                 'noise_magnitudes')}
    return group.get_controls(feats)

  n, w = _both(run)
  out = dict(seed=seed, input_checksum=checksum(inp))
  for tag, o in (('f32', n), ('wide', w)):
    out['harmonic_' + tag] = o['harmonic']['signal']
    out['filtered_noise_' + tag] = o['filtered_noise']['signal']
    out['audio_' + tag] = o['out']['signal']
  out['magnitudes'] = n['filtered_noise']['controls']['magnitudes']
  out['harmonic_distribution'] = n['harmonic']['controls']['harmonic_distribution']
  out['amplitudes'] = n['harmonic']['controls']['amplitudes']
  return out


def c2_item():
  """One batch item at the configs[1..4] shapes (F=1000, K=100, 65 bands, 64000
  samples): the reference's harmonic and filtered-noise signals, wide; stored as
  float32 (rounding 6e-8, far inside the 1e-4 gate) to keep the file small."""
  ddsp = ref_on_shim.load()
  tf = ref_on_shim.tf()
  seed = 303
  inp = synth_inputs(1, 1000, 100, 65, 64000, seed=seed)
  assert _nyquist_margin(ddsp, inp, 64000, 100) > 1e-2

  def run():
    harm = ddsp.synths.Harmonic(n_samples=64000, sample_rate=16000)
    noise = ddsp.synths.FilteredNoise(n_samples=64000, window_size=0)
    tf.random.inject_uniform(inp['noise'])
    return {'h': harm(inp['amps'], inp['harmonic_distribution'], inp['f0_hz']),
            'n': noise(inp['noise_magnitudes'])}

  n, w = _both(run)
  return dict(seed=seed, input_checksum=checksum(inp),
              harmonic_wide=w['h'].astype(np.float32),
              filtered_noise_wide=w['n'].astype(np.float32),
              # the reference's own float32 result, decimated: documents its
              # phase-accumulation error at 64000 samples
              harmonic_f32_every16=n['h'][:, ::16].astype(np.float32))


def harmonic_shifts():
  """core.harmonic_synthesis with harmonic_shifts (core.py:1084-1093), 'linear'
  and 'window' amplitudes, B=2, F=50, K=20, N=3200."""
  ddsp = ref_on_shim.load()
  seed = 404
  B, F, K, N = 2, 50, 20, 3200
  inp = synth_inputs(B, F, K, 65, N, seed=seed, f0_lo=100.0, f0_hi=500.0)
  rng = np.random.default_rng(seed)
  shifts = (0.02 * rng.standard_normal((B, F, K))).astype(np.float32)
  amps = np.abs(inp['amps']) * 0.5
  hd = np.abs(inp['harmonic_distribution'])
  hd = (hd / hd.sum(-1, keepdims=True)).astype(np.float32)
  out = dict(seed=seed, shifts=shifts, amplitudes=amps.astype(np.float32),
             harmonic_distribution=hd, f0_hz=inp['f0_hz'])
  for method in ('window', 'linear'):
    n, w = _both(lambda: ddsp.core.harmonic_synthesis(
        inp['f0_hz'], amps, harmonic_shifts=shifts, harmonic_distribution=hd,
        n_samples=N, sample_rate=16000, amp_resample_method=method))
    out['audio_f32_' + method] = n
    out['audio_wide_' + method] = w.astype(np.float64)
  return out


def resample_methods():
  """core.resample (core.py:573-642) for every method, both add_endpoint values,
  3-D and 4-D inputs, up- and down-sampling."""
  ddsp = ref_on_shim.load()
  rng = np.random.default_rng(505)
  x3 = rng.standard_normal((2, 10, 3)).astype(np.float32)
  x4 = rng.standard_normal((2, 10, 4, 3)).astype(np.float32)
  out = dict(x3=x3, x4=x4)
  for method in ('nearest', 'linear', 'cubic', 'window'):
    for ep in (True, False):
      n_up = 90 if not ep else 80      # divisible by 9 intervals / 10 frames
      n, w = _both(lambda: ddsp.core.resample(x3, n_up, method=method, add_endpoint=ep))
      out['up3_%s_%d' % (method, ep)] = n
      out['up3w_%s_%d' % (method, ep)] = w
      if method != 'window':
        out['down3_%s_%d' % (method, ep)] = _both(
            lambda: ddsp.core.resample(x3, 4, method=method, add_endpoint=ep))[0]
        out['up4_%s_%d' % (method, ep)] = _both(
            lambda: ddsp.core.resample(x4, 37, method=method, add_endpoint=ep))[0]
  return out


def angular_cumsum():
  """core.angular_cumsum (core.py:799-866) and tf.cumsum on the same angular
  frequencies, float32: [2, 2500, 3] (so the 1000-sample chunking pads)."""
  ddsp = ref_on_shim.load()
  tf = ref_on_shim.tf()
  rng = np.random.default_rng(606)
  omega = (2 * np.pi * rng.uniform(50.0, 4000.0, (2, 1, 3)) / 16000.0 *
           (1 + 0.01 * rng.standard_normal((2, 2500, 3)))).astype(np.float32)
  n, w = _both(lambda: ddsp.core.angular_cumsum(tf.convert_to_tensor(omega)))
  return dict(omega=omega, phase_f32=n, phase_wide=w.astype(np.float64))


def spectral_loss():
  """losses.SpectralLoss (losses.py:130-243) with the ae.gin weights (L1 on
  magnitudes and log magnitudes, ae.gin:39-41), B=2, N=8000."""
  ddsp = ref_on_shim.load()
  rng = np.random.default_rng(707)
  target = (0.1 * rng.standard_normal((2, 8000))).astype(np.float32)
  t = np.arange(8000) / 16000.0
  audio = (0.3 * np.sin(2 * np.pi * 220.0 * t)[None, :] * np.array([[1.0], [0.5]])
           + 0.05 * rng.standard_normal((2, 8000))).astype(np.float32)
  out = dict(target=target, audio=audio)
  for tag, kw in (('mag', dict(mag_weight=1.0, logmag_weight=0.0)),
                  ('maglog', dict(mag_weight=1.0, logmag_weight=1.0))):
    n, w = _both(lambda: ddsp.losses.SpectralLoss(**kw)(target, audio))
    out['loss_f32_' + tag] = np.float32(n)
    out['loss_wide_' + tag] = np.float64(w)
  return out


IR_CASES = [(65, 0), (65, 257), (65, 63), (65, 64), (100, 51), (100, 50), (513, 257),
            (513, 22), (1025, 257), (16, 257), (65, 3)]


def impulse_responses():
  """core.frequency_impulse_response (core.py:1534-1565) and frequency_filter
  (1628-1655) for even AND odd window sizes: tf.signal.hann_window is periodic
  for even lengths and symmetric for odd ones (window_ops._raised_cosine_window),
  which a restatement gets wrong unless it is checked on an odd window shorter than
  the impulse response (window_size=257 with more than 129 bins, e.g.)."""
  ddsp = ref_on_shim.load()
  rng = np.random.default_rng(808)
  out = {}
  for nb, ws in IR_CASES:
    m = rng.uniform(0.0, 1.0, (2, 3, nb)).astype(np.float32)
    n, w = _both(lambda: ddsp.core.frequency_impulse_response(m, ws))
    out['mags_%d_%d' % (nb, ws)] = m
    out['ir_f32_%d_%d' % (nb, ws)] = n
    out['ir_wide_%d_%d' % (nb, ws)] = w.astype(np.float64)
  noise = rng.uniform(-1.0, 1.0, (2, 960)).astype(np.float32)
  mags = rng.uniform(0.0, 1.0, (2, 20, 513)).astype(np.float32)
  n, w = _both(lambda: ddsp.core.frequency_filter(noise, mags, window_size=257))
  out.update(filter_noise=noise, filter_mags=mags, filter_f32=n,
             filter_wide=w.astype(np.float64))
  return out


FIXTURES = dict(c1_harmonic=c1_harmonic, decoder_small=decoder_small, c2_item=c2_item,
                harmonic_shifts=harmonic_shifts, resample_methods=resample_methods,
                angular_cumsum=angular_cumsum, spectral_loss=spectral_loss,
                impulse_responses=impulse_responses)


def compare(name, got, want, atol=0.0):
  assert set(want.files) == set(got), (name, sorted(set(want.files) ^ set(got)))
  for k in want.files:
    np.testing.assert_allclose(np.asarray(got[k], np.float64),
                               np.asarray(want[k], np.float64), rtol=0, atol=atol,
                               err_msg='%s/%s' % (name, k))


if __name__ == '__main__':
  check = '--check' in sys.argv
  for name, fn in FIXTURES.items():
    path = os.path.join(HERE, name + '.npz')
    got = fn()
    if check:
      compare(name, got, np.load(path))
      print('ok   ', name)
    else:
      np.savez_compressed(path, **got)
      print('wrote', name, '%.0f kB' % (os.path.getsize(path) / 1e3))
