"""Multi-GPU plumbing: one process per GPU.

* inference: contiguous block sharding of the instance dimension, replicated weights, NO collective;
* training: data-parallel REINFORCE -- one flat fp32 bucket all-reduce of the actor gradients (4.2 MB, NCCL
  over NVLink on the GPU box, gloo in the CPU tests) plus a scalar all-reduce so the EMA baseline uses the
  global reward mean (trainPNLow.py:81-86); clipping happens after the all-reduce on identical data, so
  replicas stay bit-identical.
Everything degrades to a no-op when torch.distributed is not initialised (single GPU).
"""
from __future__ import annotations

from typing import Iterable, Tuple

import torch
import torch.distributed as dist


def _on() -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def rank() -> int:
    return dist.get_rank() if _on() else 0


def world_size() -> int:
    return dist.get_world_size() if _on() else 1


def shard_range(n: int, r: int = None, w: int = None) -> Tuple[int, int]:
    """Instances [lo, hi) of rank r: ceil(n/w)-sized contiguous blocks."""
    r = rank() if r is None else r
    w = world_size() if w is None else w
    per = -(-n // w)
    return min(r * per, n), min((r + 1) * per, n)


def shard(batch: torch.Tensor) -> torch.Tensor:
    lo, hi = shard_range(batch.shape[0])
    return batch[lo:hi] if _on() else batch


def global_mean(x: torch.Tensor) -> torch.Tensor:
    """Mean over the instances of ALL ranks (sum and count all-reduced, so ragged shards are weighted right)."""
    s = torch.stack([x.sum().float(), torch.tensor(float(x.numel()), device=x.device)])
    if _on():
        dist.all_reduce(s)
    return s[0] / s[1]


def allreduce_gradients(params: Iterable[torch.nn.Parameter]) -> None:
    """Average gradients across ranks through one flat bucket (latency-bound message: one collective)."""
    if not _on():
        return
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat)
    flat /= world_size()
    off = 0
    for g in grads:
        g.copy_(flat[off:off + g.numel()].view_as(g))
        off += g.numel()


def gather_concat(x: torch.Tensor) -> torch.Tensor:
    """Reporting only: concatenate per-rank results in rank order (equal shard sizes)."""
    if not _on():
        return x
    out = [torch.empty_like(x) for _ in range(world_size())]
    dist.all_gather(out, x)
    return torch.cat(out)
