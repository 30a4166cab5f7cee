"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink on the B200 box, gloo in CPU tests).

The reference has NO distributed code: training is a single-GPU loop.  Data parallelism is therefore
defined here as the mean of the per-rank gradients of the reference step (each rank: its own 128^3 crop, batch 1,
Dice ``batch=True`` evaluated per rank).  Only the path that really has an exchange step uses a collective:

  * training  — ``DistributedDataParallel``: the hand-scheduled backward (autograd.py) writes parameter gradients
    into ONE flat fp32 buffer in completion order and reports progress; every time a bucket is complete it is
    all-reduced (SUM) on a side stream while the rest of the backward keeps running.  The 1/world mean is folded
    into the fused optimizer step (``Ranger2020.grad_scale``), so there is no separate scaling pass.
  * inference — volumes (or ensemble members) are sharded round-robin across ranks with no collective at all
    (``shard_indices``); every rank holds all models.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist
from torch import nn


def shard_indices(n_items: int, rank: int, world: int) -> List[int]:
    """Round-robin partition used for the validation cohort (BASELINE config 5): item i goes to rank i % world."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world of {world}")
    return list(range(rank, n_items, world))


class BucketReducer:
    """All-reduces consecutive slices of a flat gradient buffer as soon as they are final."""

    def __init__(self, flat: torch.Tensor, bucket_elems: int, group=None):
        self.flat, self.group = flat, group
        n = flat.numel()
        bucket_elems = max(int(bucket_elems), 1)
        self.bounds = [(b, min(b + bucket_elems, n)) for b in range(0, n, bucket_elems)]
        self.next = 0
        self.cuda = flat.is_cuda
        self.comm = torch.cuda.Stream(device=flat.device) if self.cuda else None
        self.launched = 0

    def start(self):
        self.next = 0

    def _launch(self, lo: int, hi: int, after=()):
        if self.cuda:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.flat.device))
            with torch.cuda.stream(self.comm):
                self.comm.wait_event(ev)
                for e in after:  # e.g. the weight-gradient side stream of the backward
                    self.comm.wait_event(e)
                dist.all_reduce(self.flat[lo:hi], op=dist.ReduceOp.SUM, group=self.group)
        else:
            dist.all_reduce(self.flat[lo:hi], op=dist.ReduceOp.SUM, group=self.group)
        self.launched += 1

    def progress(self, final_upto: int, after=()):
        """Gradients in flat[0:final_upto] are final once the work already enqueued on the current stream (and the
        CUDA events in ``after``) has completed."""
        while self.next < len(self.bounds) and self.bounds[self.next][1] <= final_upto:
            self._launch(*self.bounds[self.next], after=after)
            self.next += 1

    def finish(self):
        self.progress(self.flat.numel())
        if self.cuda:
            torch.cuda.current_stream(self.flat.device).wait_stream(self.comm)


class DistributedDataParallel(nn.Module):
    """Data-parallel wrapper of a brats21_b200 network (see module docstring).  ``optimizer.grad_scale`` must be
    set to ``1 / world_size`` (``attach_optimizer`` does it)."""

    def __init__(self, module: nn.Module, process_group=None, bucket_cap_mb: float = 16.0, broadcast: bool = True):
        super().__init__()
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        self.module = module
        self.group = process_group
        self.world = dist.get_world_size(process_group)
        self.bucket_elems = int(bucket_cap_mb * (1 << 20) / 4)
        self._reducer: Optional[BucketReducer] = None
        if broadcast:
            with torch.no_grad():
                for t in list(module.parameters()) + list(module.buffers()):
                    dist.broadcast(t.data, src=0, group=process_group)

    def _ensure_reducer(self):
        gs = self.module.grad_store()
        if self._reducer is None or self._reducer.flat.data_ptr() != gs.flat.data_ptr():
            self._reducer = BucketReducer(gs.flat, self.bucket_elems, self.group)
            gs.on_begin = self._reducer.start
            gs.on_ready = self._reducer.progress
            gs.on_finish = self._reducer.finish
        return self._reducer

    def attach_optimizer(self, optimizer):
        optimizer.grad_scale = 1.0 / self.world
        return optimizer

    def forward(self, *args, **kwargs):
        if self.module.training and torch.is_grad_enabled():
            self._ensure_reducer()
        return self.module(*args, **kwargs)

    def state_dict(self, *args, **kwargs):  # checkpoints stay loadable by the bare network / the reference
        return self.module.state_dict(*args, **kwargs)

    def load_state_dict(self, *args, **kwargs):
        return self.module.load_state_dict(*args, **kwargs)
