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


def _collective_device() -> torch.device:
    """Tensors handed to a collective must live where the backend works: CUDA for NCCL, host for gloo."""
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


def global_mean_count(x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """(mean, count) over the instances of ALL ranks: sum and count are all-reduced, so ragged -- even empty -- shards
    are weighted by their size."""
    s = torch.stack([x.sum().float(), torch.tensor(float(x.numel()), device=x.device)])
    if _on():
        dist.all_reduce(s)
    return s[0] / s[1], s[1]


def global_mean(x: torch.Tensor) -> torch.Tensor:
    return global_mean_count(x)[0]


def broadcast_parameters(params: Iterable[torch.Tensor], src: int = 0) -> None:
    """Make every replica start from rank ``src``'s weights (one flat broadcast)."""
    if not _on():
        return
    ps = [p for p in params]
    if not ps:
        return
    with torch.no_grad():
        flat = torch.cat([p.detach().reshape(-1) for p in ps])
        dist.broadcast(flat, src)
        off = 0
        for p in ps:
            p.copy_(flat[off:off + p.numel()].view_as(p))
            off += p.numel()


def shared_seed(src: int = 0) -> int:
    """One random 63-bit seed, drawn on rank ``src`` and broadcast: every rank shuffles the data set identically and
    ``shard`` then hands each rank its own slice of the SAME permutation."""
    t = torch.randint(0, 2 ** 62, (1,), dtype=torch.int64)
    if _on():
        t = t.to(_collective_device())
        dist.broadcast(t, src)
    return int(t.item())


def allreduce_gradients(params: Iterable[torch.nn.Parameter], average: bool = True) -> None:
    """Combine gradients across ranks through one flat bucket (latency-bound message: one collective).  Parameters
    without a gradient contribute zeros (a rank whose shard was empty still takes part).  ``average=False`` sums --
    for losses already normalised by the GLOBAL instance count (ragged shards weighted right)."""
    if not _on():
        return
    ps = [p for p in params if p.requires_grad]
    if not ps:
        return
    for p in ps:
        if p.grad is None:
            p.grad = torch.zeros_like(p)
    flat = torch.cat([p.grad.reshape(-1) for p in ps])
    dist.all_reduce(flat)
    if average:
        flat /= world_size()
    off = 0
    for p in ps:
        p.grad.copy_(flat[off:off + p.numel()].view_as(p.grad))
        off += p.numel()


def gather_concat(x: torch.Tensor) -> torch.Tensor:
    """Reporting only: concatenate per-rank results in rank order (equal shard sizes)."""
    if not _on():
        return x
    out = [torch.empty_like(x) for _ in range(world_size())]
    dist.all_gather(out, x)
    return torch.cat(out)
