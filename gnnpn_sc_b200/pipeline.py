"""Host-side driver of the greedy ML+2PN decode -- the validation loop of ``trainPNHigh.py:131-144``
(PNLow greedy -> latent -> PNHigh greedy -> actions / reward) fed from HOST batches.

The reference runs that loop batch by batch with a blocking ``.cuda()`` per batch and ``.cpu()`` per step
(trainPNHigh.py:135-144).  Here the upload of batch i+1 runs on a side stream while the kernels of batch i
execute, and the small results (picks ``int32 [K, n]``, reward ``fp32 [n]``) come back through pinned buffers.
All arithmetic stays in the CUDA kernels behind ``modelPN.CombinatorialRL`` -- this file is plumbing.
"""
from __future__ import annotations

from typing import Iterable, Iterator, Tuple

import torch


class GreedyLowHigh:
    def __init__(self, low, high, device=None):
        self.low, self.high = low.eval(), high.eval()
        self.device = torch.device(device if device is not None else "cuda")
        if self.device.type != "cuda":
            raise RuntimeError("GreedyLowHigh needs a CUDA device: the B200 path has no CPU fallback")
        self._copy = torch.cuda.Stream(self.device)
        self._stage = [None, None]            # device staging buffers (double buffer)
        self._ready = [torch.cuda.Event(), torch.cuda.Event()]
        self._consumed = [None, None]

    def _upload(self, slot: int, x_host: torch.Tensor) -> None:
        buf = self._stage[slot]
        if buf is None or buf.shape != x_host.shape:
            buf = self._stage[slot] = torch.empty(x_host.shape, device=self.device, dtype=torch.float32)
        if not x_host.is_pinned():
            x_host = x_host.pin_memory()
        with torch.cuda.stream(self._copy):
            if self._consumed[slot] is not None:          # the kernels that last read this buffer are done
                self._copy.wait_event(self._consumed[slot])
            buf.copy_(x_host, non_blocking=True)
            self._ready[slot].record(self._copy)

    @torch.no_grad()
    def run(self, host_batches: Iterable[torch.Tensor]) -> Iterator[Tuple[torch.Tensor, torch.Tensor]]:
        """Yields ``(idx_high int32 [K, n] (cpu), reward fp32 [n] (cpu))`` per host batch ``fp32 [n, L, F]``."""
        it = iter(host_batches)
        nxt = next(it, None)
        slot = 0
        if nxt is not None:
            self._upload(slot, nxt)
        main = torch.cuda.current_stream(self.device)
        while nxt is not None:
            cur_slot = slot
            nxt = next(it, None)
            if nxt is not None:                           # upload of the next batch overlaps this batch's kernels
                self._upload(cur_slot ^ 1, nxt)
            main.wait_event(self._ready[cur_slot])
            x = self._stage[cur_slot]
            _, _, _, _, latent = self.low(x, None, sample="greedy", training="SL")
            R, _, _, idx, _ = self.high(x, None, latent, sample="greedy", training="RL")
            idx32 = torch.stack(idx).to(torch.int32)
            ev = torch.cuda.Event()
            ev.record(main)
            self._consumed[cur_slot] = ev
            yield idx32.cpu(), R.cpu()                    # device->host reads synchronise this batch
            slot = cur_slot ^ 1
