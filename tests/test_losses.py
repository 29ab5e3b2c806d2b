"""SpectralLoss / stft (torch) against the NumPy oracle - runs on CPU."""
import numpy as np
import pytest
import torch

from oracle import ddsp_oracle as o
from ddsp_b200 import losses, spectral_ops


@pytest.mark.parametrize('size', [64, 256, 2048])
def test_stft_matches_oracle(size):
  """spectral_ops_test.py:26-41 intent: tf.signal.stft(pad_end=True) magnitudes."""
  rng = np.random.default_rng(size)
  audio = rng.standard_normal((2, 4000)).astype(np.float32)
  got = spectral_ops.compute_mag(torch.from_numpy(audio), size=size).numpy()
  want = o.stft_mag(audio, size)
  assert got.shape == want.shape == (2, -(-4000 // (size // 4)), size // 2 + 1)
  np.testing.assert_allclose(got, want, rtol=1e-3, atol=1e-3)


def test_spectral_loss_matches_oracle():
  rng = np.random.default_rng(0)
  a = rng.standard_normal((2, 8000)).astype(np.float32) * 0.1
  b = rng.standard_normal((2, 8000)).astype(np.float32) * 0.1
  loss = losses.SpectralLoss(mag_weight=1.0, logmag_weight=1.0)   # ae.gin:39-41
  got = float(loss(torch.from_numpy(a), torch.from_numpy(b)))
  want = o.spectral_loss(a, b, mag_weight=1.0, logmag_weight=1.0)
  assert abs(got - want) < 1e-3 * want
  assert float(loss(torch.from_numpy(a), torch.from_numpy(a))) == 0.0


def test_spectral_loss_is_scalar_and_differentiable():
  """losses_test.py:66-85: scalar, finite; plus a gradient reaches the audio."""
  loss_obj = losses.SpectralLoss(logmag_weight=1.0, delta_time_weight=1.0,
                                 delta_freq_weight=1.0, cumsum_freq_weight=1.0)
  a = torch.randn(3, 4000) * 0.1
  b = (torch.randn(3, 4000) * 0.1).requires_grad_(True)
  val = loss_obj(a, b)
  assert val.dim() == 0 and torch.isfinite(val)
  val.backward()
  assert b.grad is not None and torch.isfinite(b.grad).all() and b.grad.abs().sum() > 0
  assert set(loss_obj.get_losses_dict(a, b)) == {'spectral_loss'}


def test_mean_difference_types():
  t, v = torch.ones(2, 3), torch.zeros(2, 3)
  assert float(losses.mean_difference(t, v, 'L1')) == 1.0
  assert float(losses.mean_difference(t, v, 'l2')) == 1.0
  with pytest.raises(ValueError):
    losses.mean_difference(t, v, 'huber')
