"""GPU: the host-buffer decoder pipeline (ddsp_b200_decoder_forward_host) gives
exactly the audio of ProcessorGroup.__call__ on device tensors, for every chunk
count, and matches the float64 oracle within the parity tolerance."""
import numpy as np
import pytest
import torch

from oracle import ddsp_oracle as oracle
from tests.util import rel_err, synth_inputs

import ddsp_b200
from ddsp_b200 import host

pytestmark = pytest.mark.gpu
KEYS = ['amps', 'harmonic_distribution', 'f0_hz', 'noise_magnitudes']


def _group(N, seed):
  harm = ddsp_b200.Harmonic(n_samples=N)
  noise = ddsp_b200.FilteredNoise(n_samples=N, window_size=0, seed=seed)
  return ddsp_b200.ProcessorGroup(dag=[
      (harm, ['amps', 'harmonic_distribution', 'f0_hz']),
      (noise, ['noise_magnitudes']),
      (ddsp_b200.Add(), ['filtered_noise/signal', 'harmonic/signal'])])


@pytest.mark.parametrize('B,chunks', [(7, 1), (7, 3), (7, 7), (7, 16), (1, 4),
                                      (32, 8)])
def test_host_decoder_equals_device_call(B, chunks):
  F, K, nb, N = 125, 100, 65, 8000
  inp = synth_inputs(B, F, K, nb, N, seed=5 + B)
  feats = {k: inp[k] for k in KEYS}
  want = _group(N, seed=9)({k: torch.from_numpy(v).cuda() for k, v in feats.items()})
  dec = ddsp_b200.HostDecoder(_group(N, seed=9), max_batch=32, n_frames=F,
                              n_harmonics=K, n_bands=nb, n_chunks=chunks)
  got = dec({k: host.pin(v) for k, v in feats.items()})
  assert got.is_pinned() and tuple(got.shape) == (B, N)
  assert torch.equal(got, want.cpu())
  # second call: fresh Philox offset on both sides, staging buffers reused
  group = _group(N, seed=9)
  group({k: torch.from_numpy(v).cuda() for k, v in feats.items()})
  want2 = group({k: torch.from_numpy(v).cuda() for k, v in feats.items()})
  got2 = dec(feats)              # plain numpy inputs are accepted too
  assert torch.equal(got2, want2.cpu())
  assert not torch.equal(got2, got)
  dec.close()


def test_host_decoder_matches_oracle():
  B, F, K, nb, N = 3, 125, 100, 65, 8000
  inp = synth_inputs(B, F, K, nb, N, seed=21)
  dec = ddsp_b200.HostDecoder(_group(N, seed=4), B, F, K, nb, n_chunks=2)
  got = dec({k: inp[k] for k in KEYS}).numpy()
  nz = oracle.philox_uniform_noise(B, N, seed=4, offset=0)
  want = oracle.decoder(inp['amps'], inp['harmonic_distribution'], inp['f0_hz'],
                        inp['noise_magnitudes'], nz, n_samples=N, window_size=0,
                        dtype=np.float64)['add']['signal']
  emax, el2 = rel_err(got, want)
  assert emax < 1e-4 and el2 < 1e-4, (emax, el2)


def test_host_decoder_shape_errors():
  B, F, K, nb, N = 2, 125, 100, 65, 8000
  inp = synth_inputs(B, F, K, nb, N, seed=1)
  dec = ddsp_b200.HostDecoder(_group(N, seed=0), 1, F, K, nb)
  with pytest.raises(ValueError):
    dec({k: inp[k] for k in KEYS})          # batch 2 > max_batch 1
  with pytest.raises(ValueError):
    dec({k: torch.from_numpy(inp[k]).cuda() for k in KEYS})
