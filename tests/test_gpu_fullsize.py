"""Full-size (BASELINE.json shapes) GPU tests through size-independent
properties, plus oracle spot checks on a few items of the big batch.

C2: B=32, F=1000, K=100, nb=65, N=64000.  (C3's B=256 only changes the batch
axis, which the item-independence test covers.)
"""
import numpy as np
import pytest
import torch

from oracle import ddsp_oracle as oracle
from tests.util import rel_err, synth_inputs

import ddsp_b200
from ddsp_b200 import core

pytestmark = pytest.mark.gpu
B, F, K, NB, N = 32, 1000, 100, 65, 64000
TOL = 1e-4


@pytest.fixture(scope='module')
def c2():
  inp = synth_inputs(B, F, K, NB, N, seed=2024)
  dev = {k: torch.from_numpy(v).cuda() for k, v in inp.items()}
  amps, hd = core.harmonic_controls(dev['amps'], dev['harmonic_distribution'],
                                    dev['f0_hz'], 16000)
  mags = core.noise_controls(dev['noise_magnitudes'], -5.0)
  return inp, dev, amps, hd, mags


def _np(x):
  return x.detach().cpu().numpy()


def test_oracle_spot_check_full_length(c2):
  """Items 0 and 17 of the C2 batch against the float64 arbiter at N=64000."""
  inp, dev, amps, hd, mags = c2
  audio = _np(core.harmonic_synthesis(dev['f0_hz'], amps,
                                      harmonic_distribution=hd, n_samples=N))
  nz = _np(core.filtered_noise(mags, N, window_size=0, noise=dev['noise']))
  for b in (0, 17):
    sl = slice(b, b + 1)
    want = oracle.harmonic_synthesis(
        inp['f0_hz'][sl], _np(amps)[sl], harmonic_distribution=_np(hd)[sl],
        n_samples=N, dtype=np.float64)
    emax, el2 = rel_err(audio[sl], want)
    assert emax < TOL and el2 < TOL, ('harmonic', b, emax, el2)
    want_n = oracle.frequency_filter(inp['noise'][sl], _np(mags)[sl].astype(np.float64),
                                     window_size=0)
    emax, el2 = rel_err(nz[sl], want_n)
    assert emax < TOL and el2 < TOL, ('noise', b, emax, el2)


def test_constant_f0_phase_is_exact_at_full_length():
  """4 s of a 100-harmonic tone: inclusive-cumsum closed form, no drift."""
  f0 = np.full((2, F, 1), 77.7, np.float32)          # all 100 harmonics live
  f0[1] = 79.9
  amp = np.ones((2, F, 1), np.float32)
  hd = np.full((2, F, K), 1.0 / K, np.float32)
  got = _np(core.harmonic_synthesis(f0, amp, harmonic_distribution=hd, n_samples=N))
  t = (np.arange(N, dtype=np.float64) + 1) / 16000.0
  for b in range(2):
    fb = float(f0[b, 0, 0])                           # the float32 value
    k = np.arange(1, K + 1, dtype=np.float64)
    want = np.zeros(N)
    for kk in k:                                      # avoid an [N, K] temp
      want += np.sin(2 * np.pi * fb * kk * t) / K
    assert np.abs(got[b] - want).max() < 1e-4 * np.abs(want).max()


def test_harmonic_is_linear_in_amplitude(c2):
  _, dev, amps, hd, _ = c2
  a = core.harmonic_synthesis(dev['f0_hz'], amps, harmonic_distribution=hd, n_samples=N)
  b = core.harmonic_synthesis(dev['f0_hz'], amps * 2.0, harmonic_distribution=hd,
                              n_samples=N)
  emax, _ = rel_err(_np(b), 2.0 * _np(a))
  assert emax < 2e-6


def test_items_are_independent(c2):
  """Batch sharding: a shard computed alone equals its rows in the
  full batch - bit for bit (same kernel, same per-item arithmetic)."""
  _, dev, amps, hd, mags = c2
  full = core.harmonic_synthesis(dev['f0_hz'], amps, harmonic_distribution=hd, n_samples=N)
  part = core.harmonic_synthesis(dev['f0_hz'][8:16], amps[8:16],
                                 harmonic_distribution=hd[8:16], n_samples=N)
  assert torch.equal(full[8:16], part)
  fulln = core.filtered_noise(mags, N, window_size=0, noise=dev['noise'])
  partn = core.filtered_noise(mags[8:16], N, window_size=0, noise=dev['noise'][8:16])
  assert torch.equal(fulln[8:16], partn)


def test_harmonic_is_causal_in_frames(c2):
  """audio[t] depends on frames <= t // hop + 1 only: synthesising the first
  half of the frames reproduces the first half of the audio (minus one hop)."""
  _, dev, amps, hd, _ = c2
  full = core.harmonic_synthesis(dev['f0_hz'][:4], amps[:4],
                                 harmonic_distribution=hd[:4], n_samples=N)
  half = core.harmonic_synthesis(dev['f0_hz'][:4, :F // 2].contiguous(),
                                 amps[:4, :F // 2].contiguous(),
                                 harmonic_distribution=hd[:4, :F // 2].contiguous(),
                                 n_samples=N // 2)
  n_ok = N // 2 - N // F
  emax, _ = rel_err(_np(half[:, :n_ok]), _np(full[:, :n_ok]))
  assert emax < 1e-6


def test_noise_filter_is_linear_in_magnitudes(c2):
  _, dev, _, _, mags = c2
  m1, m2 = mags[:8], mags[8:16]
  nz = dev['noise'][:8]
  a = core.filtered_noise(m1, N, window_size=0, noise=nz)
  b = core.filtered_noise(m2, N, window_size=0, noise=nz)
  ab = core.filtered_noise(m1 + m2, N, window_size=0, noise=nz)
  emax, _ = rel_err(_np(ab), _np(a) + _np(b))
  assert emax < 5e-6
  a3 = core.filtered_noise(m1 * 3.0, N, window_size=0, noise=nz)
  emax, _ = rel_err(_np(a3), 3.0 * _np(a))
  assert emax < 5e-6


def test_flat_magnitudes_delay_by_two_samples(c2):
  """Flat magnitudes g: the 128-tap windowed IR is g * delta at tap 64 and the
  reference crops 62, so out[t] = g x[t-2] (SURVEY 8c known answer)."""
  _, dev, _, _, _ = c2
  g = 0.37
  mags = torch.full((4, F, NB), g, device='cuda')
  nz = dev['noise'][:4]
  out = _np(core.filtered_noise(mags, N, window_size=0, noise=nz))
  x = _np(nz)
  assert np.abs(out[:, 2:] - g * x[:, :-2]).max() < 2e-6
  assert np.abs(out[:, :2]).max() < 2e-6


def test_fused_decoder_equals_sum_of_parts(c2):
  """ProcessorGroup fused path (2 launches from raw outputs) == harmonic +
  noise computed through get_controls / get_signal, same Philox stream."""
  inp, dev, amps, hd, mags = c2
  harm = ddsp_b200.Harmonic(n_samples=N)
  noise = ddsp_b200.FilteredNoise(n_samples=N, window_size=0, seed=9)
  group = ddsp_b200.ProcessorGroup(dag=[
      (harm, ['amps', 'harmonic_distribution', 'f0_hz']),
      (noise, ['noise_magnitudes']),
      (ddsp_b200.Add(), ['filtered_noise/signal', 'harmonic/signal'])])
  feats = {k: dev[k] for k in ['amps', 'harmonic_distribution', 'f0_hz',
                               'noise_magnitudes']}
  fused = group(feats)                                    # Philox offset 0
  h = core.harmonic_synthesis(dev['f0_hz'], amps, harmonic_distribution=hd, n_samples=N)
  nzs = core.filtered_noise(mags, N, window_size=0, seed=9, offset=0)
  emax, el2 = rel_err(_np(fused), _np(h) + _np(nzs))
  assert emax < 1e-5 and el2 < 1e-5, (emax, el2)
  # successive calls draw fresh noise (synths.py:192-193 semantics)
  again = group(feats)
  assert not torch.equal(fused, again)


def test_in_kernel_noise_statistics():
  x = _np(core.uniform_noise(8, N, seed=5))
  assert x.min() >= -1.0 and x.max() < 1.0
  assert abs(x.mean()) < 5e-3 and abs(x.var() - 1.0 / 3.0) < 3e-3
  # neighbouring samples and neighbouring items are uncorrelated
  assert abs(np.mean(x[:, 1:] * x[:, :-1])) < 3e-3
  assert abs(np.mean(x[0] * x[1])) < 5e-3


# ---- configs[2] / configs[3] at their real batch sizes ------------------------
def test_c3_batch256_spot_check_against_oracle():
  """BASELINE.json configs[2] (= one GPU's share of configs[4]): B=256 from raw
  network outputs through the fused ProcessorGroup route; items 0, 101 and 255
  against the float64 arbiter (which tests/test_reference_pin.py pins to the
  reference at exactly this item shape)."""
  B3 = 256
  inp = synth_inputs(B3, F, K, NB, N, seed=3003)
  harm = ddsp_b200.Harmonic(n_samples=N)
  noise = ddsp_b200.FilteredNoise(n_samples=N, window_size=0)
  group = ddsp_b200.ProcessorGroup(dag=[
      (harm, ['amps', 'harmonic_distribution', 'f0_hz']),
      (noise, ['noise_magnitudes']),
      (ddsp_b200.Add(), ['filtered_noise/signal', 'harmonic/signal'])])
  noise.injected_noise = torch.from_numpy(inp['noise']).cuda()
  feats = {k: torch.from_numpy(inp[k]).cuda() for k in
           ('amps', 'harmonic_distribution', 'f0_hz', 'noise_magnitudes')}
  audio = _np(group(feats))
  assert audio.shape == (B3, N)
  for b in (0, 101, 255):
    sl = slice(b, b + 1)
    want = oracle.decoder(inp['amps'][sl], inp['harmonic_distribution'][sl],
                          inp['f0_hz'][sl], inp['noise_magnitudes'][sl],
                          inp['noise'][sl], n_samples=N, window_size=0,
                          dtype=np.float64)['add']['signal']
    emax, el2 = rel_err(audio[sl], want)
    assert emax < TOL and el2 < TOL, (b, emax, el2)


def _ref_spectral_loss_f64(target, audio):
  """losses.SpectralLoss (ae.gin weights) per batch item, float64 torch ops:
  tf.signal.stft(pad_end=True) framing, periodic Hann, |rfft|, L1 on magnitudes and
  on safe_log magnitudes (losses.py:194-243, spectral_ops.py:34-47)."""
  total = 0.0
  for size in (2048, 1024, 512, 256, 128, 64):
    step = size // 4
    n = audio.shape[-1]
    n_frames = -(-n // step)
    pad = (n_frames - 1) * step + size - n
    win = torch.hann_window(size, periodic=True, dtype=torch.float64, device=audio.device)

    def mag(x):
      fr = torch.nn.functional.pad(x, (0, pad)).unfold(-1, size, step)
      return torch.fft.rfft(fr * win, dim=-1).abs()

    t, v = mag(target), mag(audio)
    slog = lambda m: torch.log(torch.where(m <= 0, torch.full_like(m, 1e-5), m))
    total = total + (t - v).abs().mean(dim=(-2, -1)) + (slog(t) - slog(v)).abs().mean(dim=(-2, -1))
  return total            # [items]


def test_c4_batch128_loss_and_gradients():
  """BASELINE.json configs[3]: decoder forward + backward through the multi-scale
  SpectralLoss at B=128, against float64 autograd of an op-by-op restatement, on
  two items of the batch:
    * the audio and the loss value (1e-4);
    * the decoder's backward kernels at this shape, driven by the float64 loss
      gradient dL/d audio (so that what is compared is OUR backward, not the
      conditioning of an L1-of-log-magnitude loss in float32);
    * the whole float32 chain (loss backward included): direction of the gradient.
  """
  from ddsp_b200 import autograd as ag
  from ddsp_b200 import losses
  from tests.test_gpu_backward import ref_harmonic, ref_noise
  B4, items = 128, (5, 77)
  inp = synth_inputs(B4, F, K, NB, N, seed=4004)
  dev = torch.device('cuda')
  raw = {k: torch.from_numpy(inp[k]).to(dev).requires_grad_(True) for k in
         ('amps', 'harmonic_distribution', 'noise_magnitudes')}
  f0 = torch.from_numpy(inp['f0_hz']).to(dev)
  target = 0.1 * torch.randn(B4, N, device=dev, generator=torch.Generator(dev).manual_seed(1))
  nz = core.uniform_noise(B4, N, seed=9, offset=4)       # the in-kernel Philox stream
  loss_obj = losses.SpectralLoss(mag_weight=1.0, logmag_weight=1.0)
  audio = ag.decoder_train(raw['amps'], raw['harmonic_distribution'], f0,
                           raw['noise_magnitudes'], n_samples=N, window_size=0,
                           seed=9, offset=4)

  idx = torch.tensor(items, device=dev)
  r64 = {k: v.detach()[idx].double().requires_grad_(True) for k, v in raw.items()}
  a, h = ag.harmonic_controls(r64['amps'], r64['harmonic_distribution'], f0[idx].double())
  ref_audio = (ref_harmonic(f0[idx].double(), a, h, N) +
               ref_noise(ag.exp_sigmoid(r64['noise_magnitudes'] - 5.0), nz[idx].double(), N))
  emax, el2 = rel_err(_np(audio.detach()[idx]), _np(ref_audio.detach()))
  assert emax < TOL and el2 < TOL, (emax, el2)
  per_item = _ref_spectral_loss_f64(target[idx].double(), ref_audio)
  # value: the same two items as their own batch
  with torch.no_grad():
    sub = loss_obj(target[idx], audio.detach()[idx])
  assert abs(float(sub) - float(per_item.mean())) < 1e-4 * float(per_item.mean())

  # loss = mean over items of the per-item losses
  g_audio = torch.autograd.grad(per_item.sum() / B4, ref_audio, retain_graph=True)[0]
  (per_item.sum() / B4).backward()
  want = {k: r64[k].grad for k in raw}

  # (a) our backward kernels at B = 128, fed the float64 loss gradient
  G = torch.zeros_like(audio)
  G[idx] = g_audio.float()
  audio.backward(gradient=G, retain_graph=True)
  for k in raw:
    got = raw[k].grad[idx].double()
    err = float((got - want[k]).abs().max() / want[k].abs().max())
    l2 = float(((got - want[k])**2).sum().sqrt() / (want[k]**2).sum().sqrt())
    assert err < 2e-3 and l2 < 1e-3, (k, err, l2)
    others = raw[k].grad.clone()
    others[idx] = 0
    assert float(others.abs().max()) == 0.0        # no leakage between batch items
    raw[k].grad = None

  # (b) the whole float32 chain.  L1 of LOG magnitudes has gradient sign / |X| per
  # bin: bins where the synthesized spectrum is tiny make it ill-conditioned in
  # float32 (any float32 evaluation, TensorFlow's included), so only the direction
  # is checked here; (a) is the accuracy statement about our kernels.
  loss_obj(target, audio).backward()
  for k in raw:
    got = raw[k].grad[idx].double().flatten()
    w = want[k].flatten()
    cos = float((got * w).sum() / (got.norm() * w.norm()))
    assert cos > 0.95, (k, cos)
