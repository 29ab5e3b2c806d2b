"""Batch sharding across the GPUs of one box.

Every batch item of the decoder is independent (no op of the path reduces over
the batch), so the N-GPU path is: contiguous batch shards, one process per GPU,
NO collective on the synthesis path.  The only optional exchange is an
all-gather that reassembles the [B, N] audio on every rank (NCCL over NVLink on
GPUs; the same code runs on gloo/CPU tensors in the tests).
"""
from typing import Dict, Tuple

import torch
import torch.distributed as dist


def shard_bounds(batch: int, rank: int, world: int) -> Tuple[int, int]:
  """Items [lo, hi) of rank `rank`: contiguous, sizes differ by at most one."""
  if world < 1 or not 0 <= rank < world:
    raise ValueError(f'bad rank/world: {rank}/{world}')
  base, extra = divmod(batch, world)
  lo = rank * base + min(rank, extra)
  hi = lo + base + (1 if rank < extra else 0)
  return lo, hi


def shard_batch(inputs: Dict[str, torch.Tensor], rank: int,
                world: int) -> Dict[str, torch.Tensor]:
  """Slices every tensor of a feature dict along the batch axis."""
  sizes = {int(v.shape[0]) for v in inputs.values()}
  if len(sizes) != 1:
    raise ValueError(f'inputs disagree on the batch size: {sorted(sizes)}')
  lo, hi = shard_bounds(sizes.pop(), rank, world)
  return {k: v[lo:hi] for k, v in inputs.items()}


def all_gather_audio(audio: torch.Tensor, batch: int, group=None) -> torch.Tensor:
  """Reassembles the full [batch, N] audio on every rank (optional; off the
  synthesis path).  Shards may be ragged by one item, so pad to the max shard."""
  world = dist.get_world_size(group)
  rank = dist.get_rank(group)
  lo, hi = shard_bounds(batch, rank, world)
  if audio.shape[0] != hi - lo:
    raise ValueError(f'rank {rank} holds {audio.shape[0]} items, expected {hi - lo}')
  max_items = -(-batch // world)
  padded = audio
  if audio.shape[0] < max_items:
    pad = torch.zeros((max_items - audio.shape[0],) + tuple(audio.shape[1:]),
                      dtype=audio.dtype, device=audio.device)
    padded = torch.cat([audio, pad], 0)
  out = torch.empty((world * max_items,) + tuple(audio.shape[1:]),
                    dtype=audio.dtype, device=audio.device)
  dist.all_gather_into_tensor(out, padded.contiguous(), group=group)
  parts = []
  for r in range(world):
    rlo, rhi = shard_bounds(batch, r, world)
    parts.append(out[r * max_items:r * max_items + (rhi - rlo)])
  return torch.cat(parts, 0)


def max_over_ranks(value: float, device=None, group=None) -> float:
  """Max of a scalar over ranks (timing rule: report the slowest rank)."""
  if not (dist.is_available() and dist.is_initialized()):
    return float(value)
  t = torch.tensor([float(value)], dtype=torch.float64, device=device)
  dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
  return float(t.item())
