"""The N>1 path: contiguous batch shards, one process per GPU, no collective on
the synthesis path, optional all-gather of the audio.

CPU (gloo, world 2 and 3): the sharding / gather logic around the real decoder
arithmetic - each rank synthesizes ITS shard with the CPU oracle (a small DAG) and
the gathered audio must equal the unsharded oracle result item for item.
GPU (`-m gpu`, needs >= 2 devices): the same with the CUDA decoder on two B200s
over NCCL - the gathered audio must equal the single-GPU result bit for bit (every
shard is fed its slice of the same Philox stream).
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ddsp_b200 import sharding


def test_shard_bounds_cover_batch_exactly():
  for batch in (0, 1, 5, 32, 255, 2048):
    for world in (1, 2, 3, 8):
      spans = [sharding.shard_bounds(batch, r, world) for r in range(world)]
      assert spans[0][0] == 0 and spans[-1][1] == batch
      for (a, b), (c, d) in zip(spans, spans[1:]):
        assert b == c
      sizes = [b - a for a, b in spans]
      assert max(sizes) - min(sizes) <= 1
  with pytest.raises(ValueError):
    sharding.shard_bounds(4, 2, 2)


def test_shard_batch_checks_batch_sizes():
  with pytest.raises(ValueError):
    sharding.shard_batch({'a': torch.zeros(4, 2), 'b': torch.zeros(3, 2)}, 0, 2)


def _free_port():
  with socket.socket() as s:
    s.bind(('127.0.0.1', 0))
    return s.getsockname()[1]


_SHAPE = dict(F=10, K=12, nb=65, N=640)     # a small ae.gin decoder DAG


def _oracle_decoder(inp):
  from oracle import ddsp_oracle as o
  out = o.decoder(inp['amps'], inp['harmonic_distribution'], inp['f0_hz'],
                  inp['noise_magnitudes'], inp['noise'], n_samples=_SHAPE['N'],
                  window_size=0, dtype=np.float64)
  return out['add']['signal']


def _worker(rank, world, port, batch, q):
  os.environ['MASTER_ADDR'] = '127.0.0.1'
  os.environ['MASTER_PORT'] = str(port)
  dist.init_process_group('gloo', rank=rank, world_size=world)
  try:
    from tests.util import synth_inputs
    inp = synth_inputs(batch, _SHAPE['F'], _SHAPE['K'], _SHAPE['nb'], _SHAPE['N'], seed=5)
    full = {k: torch.from_numpy(v) for k, v in inp.items()}
    mine = sharding.shard_batch(full, rank, world)
    lo, hi = sharding.shard_bounds(batch, rank, world)
    assert mine['f0_hz'].shape[0] == hi - lo
    # the synthesis of this rank's shard: per item, no communication
    audio = torch.from_numpy(_oracle_decoder({k: v.numpy() for k, v in mine.items()})
                             ) if hi > lo else torch.zeros((0, _SHAPE['N']), dtype=torch.float64)
    gathered = sharding.all_gather_audio(audio, batch)
    want = torch.from_numpy(_oracle_decoder(inp))
    assert gathered.shape == want.shape
    assert torch.equal(gathered, want)
    slowest = sharding.max_over_ranks(10.0 + rank)
    assert slowest == 10.0 + world - 1
    q.put((rank, 'ok'))
  except Exception as e:  # pylint: disable=broad-except
    q.put((rank, repr(e)))
  finally:
    dist.destroy_process_group()


@pytest.mark.parametrize('world,batch', [(2, 8), (2, 7), (3, 10)])
def test_sharded_replicas_gloo(world, batch):
  ctx = mp.get_context('spawn')
  q = ctx.Queue()
  port = _free_port()
  procs = [ctx.Process(target=_worker, args=(r, world, port, batch, q))
           for r in range(world)]
  for p in procs:
    p.start()
  results = [q.get(timeout=120) for _ in procs]
  for p in procs:
    p.join(timeout=60)
  assert sorted(results) == [(r, 'ok') for r in range(world)], results


# ---- two GPUs: the CUDA decoder on NCCL ---------------------------------------
def _gpu_worker(rank, world, port, batch, q):
  os.environ['MASTER_ADDR'] = '127.0.0.1'
  os.environ['MASTER_PORT'] = str(port)
  torch.cuda.set_device(rank)
  dist.init_process_group('nccl', rank=rank, world_size=world,
                          device_id=torch.device('cuda', rank))
  try:
    import ddsp_b200
    from ddsp_b200 import core
    from tests.util import synth_inputs
    F, K, nb, N = 250, 100, 65, 16000
    inp = synth_inputs(batch, F, K, nb, N, seed=6)
    keys = ('amps', 'harmonic_distribution', 'f0_hz', 'noise_magnitudes')
    full = {k: torch.from_numpy(inp[k]).cuda() for k in keys}
    lo, hi = sharding.shard_bounds(batch, rank, world)
    mine = sharding.shard_batch(full, rank, world)
    # decoder_forward keys the in-kernel Philox stream by the item index inside the
    # call; a shard passes its global offset through the injected stream instead
    noise_full = core.uniform_noise(batch, N, seed=11, offset=3)
    audio = core.decoder_forward(mine['amps'], mine['harmonic_distribution'],
                                 mine['f0_hz'], mine['noise_magnitudes'], N,
                                 window_size=0, noise=noise_full[lo:hi].contiguous())
    gathered = sharding.all_gather_audio(audio, batch)
    want = core.decoder_forward(full['amps'], full['harmonic_distribution'],
                                full['f0_hz'], full['noise_magnitudes'], N,
                                window_size=0, seed=11, offset=3)
    torch.cuda.synchronize()
    assert gathered.shape == (batch, N)
    assert torch.equal(gathered, want), float((gathered - want).abs().max())
    assert sharding.max_over_ranks(1.0 + rank, device='cuda') == float(world)
    q.put((rank, 'ok'))
  except Exception as e:  # pylint: disable=broad-except
    q.put((rank, repr(e)))
  finally:
    dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.parametrize('batch', [8, 7])
def test_sharded_decoder_two_gpus_nccl(batch):
  if torch.cuda.device_count() < 2:
    pytest.skip('needs two GPUs (gpurun --gpus 2)')
  ctx = mp.get_context('spawn')
  q = ctx.Queue()
  port = _free_port()
  procs = [ctx.Process(target=_gpu_worker, args=(r, 2, port, batch, q)) for r in range(2)]
  for p in procs:
    p.start()
  results = [q.get(timeout=300) for _ in procs]
  for p in procs:
    p.join(timeout=60)
  assert sorted(results) == [(0, 'ok'), (1, 'ok')], results
