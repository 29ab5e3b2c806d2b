"""N>1 host logic on CPU: gloo, world_size 2 and 3 (no GPU needed)."""
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


def _worker(rank, world, port, batch, q):
  os.environ['MASTER_ADDR'] = '127.0.0.1'
  os.environ['MASTER_PORT'] = str(port)
  dist.init_process_group('gloo', rank=rank, world_size=world)
  try:
    full = {'x': torch.arange(batch * 3, dtype=torch.float32).reshape(batch, 3),
            'y': torch.arange(batch, dtype=torch.float32).reshape(batch, 1)}
    mine = sharding.shard_batch(full, rank, world)
    lo, hi = sharding.shard_bounds(batch, rank, world)
    assert mine['x'].shape[0] == hi - lo
    # the "synthesis": a per-item function, no communication
    audio = mine['x'] * 2.0 + mine['y']
    gathered = sharding.all_gather_audio(audio, batch)
    want = full['x'] * 2.0 + full['y']
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
