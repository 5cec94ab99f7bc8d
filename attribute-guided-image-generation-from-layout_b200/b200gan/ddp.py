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

import os

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


_GRADUATED = os.environ.get("B200_DDP_GRADUATED", "1") != "0"


class GradBucketer:
    """Bucketed, overlapped gradient all-reduce (mean).

    Each bucket owns ONE persistent flat fp32 buffer.  When the last gradient of a bucket lands (post-accumulate-grad
    hook) the bucket's gradients are gathered into the buffer by a single multi-tensor copy and the all-reduce is launched
    asynchronously (NCCL: ReduceOp.AVG — the 1/world scale happens inside the collective); `finish()` waits and re-points
    every `p.grad` at its slice of the reduced buffer — no scale pass, no copy back (round 1 paid cat + div + copy back,
    ~1 GB of extra traffic per step).  Buckets are filled in reverse parameter order (the order backward produces
    gradients); the LAST bucket to complete (the first layers of the network) is the only one whose all-reduce cannot hide
    behind backward compute, so it is kept small (`tail_bytes`)."""

    def __init__(self, params: List[torch.nn.Parameter], bucket_bytes: int = 25 << 20, group=None,
                 tail_bytes: int = 2 << 20):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.params = [p for p in params if p.requires_grad]
        backend = dist.get_backend(group) if dist.is_initialized() else "none"
        self.avg_in_collective = backend == "nccl"
        # reverse parameter order; the tail (first parameters) forms its own small bucket
        order = list(reversed(self.params))
        tail, tsize = [], 0
        while order and tsize + order[-1].numel() * 4 <= tail_bytes:
            tail.insert(0, order.pop())
            tsize += tail[0].numel() * 4
        self.buckets: List[List[torch.nn.Parameter]] = []
        remaining = sum(p.numel() * p.element_size() for p in order)
        cur, size = [], 0
        for p in order:
            cur.append(p)
            nb = p.numel() * p.element_size()
            size += nb
            remaining -= nb
            # graduated sizes: the buckets that complete near the END of backward have little compute left to hide behind, so
            # the last ~bucket_bytes worth of gradients goes out in quarter-size buckets
            target = bucket_bytes if (remaining > bucket_bytes or not _GRADUATED) else max(bucket_bytes // 4, 1 << 20)
            if size >= target:
                self.buckets.append(cur)
                cur, size = [], 0
        if cur:
            self.buckets.append(cur)
        if tail:
            self.buckets.append(tail)
        self.bucket_of = {}
        self.flat: List[torch.Tensor] = []
        self.views: List[List[torch.Tensor]] = []
        for bi, bk in enumerate(self.buckets):
            n = sum(p.numel() for p in bk)
            flat = torch.zeros((n,), dtype=torch.float32, device=bk[0].device)
            views, off = [], 0
            for p in bk:
                self.bucket_of[p] = bi
                views.append(flat[off:off + p.numel()].view_as(p))
                off += p.numel()
            self.flat.append(flat)
            self.views.append(views)
        self.enabled = False
        self._pending = [0] * len(self.buckets)
        self._work: List[Optional[tuple]] = [None] * len(self.buckets)
        self._handles = [p.register_post_accumulate_grad_hook(self._hook) for p in self.params]

    def bucket_bytes(self) -> List[int]:
        return [f.numel() * 4 for f in self.flat]

    def arm(self):
        """Call right before backward(): gradients produced from now on are reduced."""
        self.enabled = True
        self._pending = [len(b) for b in self.buckets]
        self._work = [None] * len(self.buckets)
        # the stream backward() is called from: gradients of sub-graphs that ran on it (not on a forked stream) complete there
        self._ambient = torch.cuda.current_stream(self.flat[0].device) if (self.flat and self.flat[0].is_cuda) else None

    def _launch(self, bi: int):
        bucket, views, flat = self.buckets[bi], self.views[bi], self.flat[bi]
        if flat.is_cuda:
            # the bucket's gradients may have been produced on forked streams (the three discriminators run side by side,
            # TrainStep._side_by_side): the copy and the collective are ordered after everything those streams hold so far
            from . import ops
            cur = torch.cuda.current_stream(flat.device)
            for st in ops.all_forked_streams(flat.device) + ([self._ambient] if getattr(self, "_ambient", None) is not None else []):
                if st != cur:
                    cur.wait_stream(st)
        src = []
        missing = []
        for p, v in zip(bucket, views):
            if p.grad is None:                     # parameter without a gradient this step contributes zeros ...
                v.zero_()
                src.append(v)
                missing.append(p)
            else:
                src.append(p.grad)
        if any(s is not v for s, v in zip(src, views)):
            torch._foreach_copy_([v for s, v in zip(src, views) if s is not v], [s for s, v in zip(src, views) if s is not v])
        work = None
        if self.world > 1:
            op = dist.ReduceOp.AVG if self.avg_in_collective else dist.ReduceOp.SUM
            work = dist.all_reduce(flat, op=op, group=self.group, async_op=True)
        self._work[bi] = (work, missing)

    def _hook(self, p):
        if not self.enabled:
            return
        bi = self.bucket_of[p]
        self._pending[bi] -= 1
        if self._pending[bi] == 0:
            self._launch(bi)

    def finish(self):
        """Drain: launch buckets whose parameters did not all receive a gradient, wait, and hand every parameter its slice of
        the reduced buffer as `.grad`."""
        if not self.enabled:
            return
        for bi in range(len(self.buckets)):
            if self._work[bi] is None and self._pending[bi] > 0:
                self._launch(bi)
        for bi, item in enumerate(self._work):
            if item is None:
                continue
            work, missing = item
            if work is not None:
                work.wait()
            if self.world > 1 and not self.avg_in_collective:
                self.flat[bi].div_(self.world)
            for p, v in zip(self.buckets[bi], self.views[bi]):
                # ... and keeps `.grad is None` (optimizers skip it, as without data parallelism)
                p.grad = None if any(p is q for q in missing) else v
        self.enabled = False

    def remove(self):
        for h in self._handles:
            h.remove()
