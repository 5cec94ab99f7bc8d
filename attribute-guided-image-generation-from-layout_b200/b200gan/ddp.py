"""Data-parallel plumbing: one process per GPU, batch sharded by image, bucketed gradient all-reduce (mean) over
torch.distributed (NCCL over NVLink 5 / NVSwitch on the GPU box, gloo in the CPU tests) overlapped with backward.

The reference has no distributed code on its path; the only exchange a data-parallel G+D step needs
is one all-reduce of parameter gradients per backward.  Buckets are filled in reverse parameter order
(the order autograd produces gradients), launched asynchronously from post-accumulate-grad hooks as soon as the last
gradient of a bucket lands, and drained before the optimizer step.  Batch-norm statistics stay per shard (DDP
semantics); spectral-norm u/v evolve identically on every rank because they depend on the weights only.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist


def broadcast_module(module: torch.nn.Module, src: int = 0, group=None):
    """Make every rank start from rank `src`'s parameters and buffers."""
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src=src, group=group)
    from . import ops
    ops.bump_weight_epoch()        # `.data` writes do not move tensor versions: invalidate the packed GEMM operands


def shard_images(n_images: int, rank: int, world: int):
    """Contiguous image range of this rank (objects travel with their image)."""
    per = (n_images + world - 1) // world
    lo = min(n_images, rank * per)
    return lo, min(n_images, lo + per)


class GradBucketer:
    def __init__(self, params: List[torch.nn.Parameter], bucket_bytes: int = 25 << 20, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.params = [p for p in params if p.requires_grad]
        self.buckets: List[List[torch.nn.Parameter]] = []
        cur, size = [], 0
        for p in reversed(self.params):
            cur.append(p)
            size += p.numel() * p.element_size()
            if size >= bucket_bytes:
                self.buckets.append(cur)
                cur, size = [], 0
        if cur:
            self.buckets.append(cur)
        self.bucket_of = {}
        for bi, bk in enumerate(self.buckets):
            for p in bk:
                self.bucket_of[p] = bi
        self.enabled = False
        self._pending = [0] * len(self.buckets)
        self._work: List[Optional[tuple]] = [None] * len(self.buckets)
        self._handles = [p.register_post_accumulate_grad_hook(self._hook) for p in self.params]

    def arm(self):
        """Call right before backward(): gradients produced from now on are reduced."""
        self.enabled = True
        self._pending = [len(b) for b in self.buckets]
        self._work = [None] * len(self.buckets)

    def _launch(self, bi: int):
        bucket = [p for p in self.buckets[bi] if p.grad is not None]
        if not bucket:
            return
        flat = torch.cat([p.grad.reshape(-1) for p in bucket])
        work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True) if self.world > 1 else None
        self._work[bi] = (flat, work, bucket)

    def _hook(self, p):
        if not self.enabled:
            return
        bi = self.bucket_of[p]
        self._pending[bi] -= 1
        if self._pending[bi] == 0:
            self._launch(bi)

    def finish(self):
        """Drain: launch buckets whose parameters did not all receive a gradient, wait, average, scatter back."""
        if not self.enabled:
            return
        for bi in range(len(self.buckets)):
            if self._work[bi] is None and self._pending[bi] > 0:
                self._launch(bi)
        for item in self._work:
            if item is None:
                continue
            flat, work, bucket = item
            if work is not None:
                work.wait()
            if self.world > 1:
                flat.div_(self.world)
            off = 0
            views = []
            for p in bucket:
                n = p.numel()
                views.append(flat[off:off + n].view_as(p))
                off += n
            torch._foreach_copy_([p.grad for p in bucket], views)
        self.enabled = False

    def remove(self):
        for h in self._handles:
            h.remove()
