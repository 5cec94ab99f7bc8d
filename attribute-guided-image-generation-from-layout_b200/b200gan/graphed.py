"""CUDA-graph replay of the training iteration, one captured graph per batch LAYOUT.

A layout = the per-image object counts (train64.py:141: the loader yields 3-9 objects per image, data/vg_custom_mask.py:45):
it fixes every tensor shape of the iteration and the host-built index plans (ConvLSTM time-major packing, crop grouping),
so an iteration can be replayed from a captured graph only for the layout it was captured with.  `GraphedTrainStep` keeps an
LRU cache of captured iterations keyed by layout: a batch whose layout is cached is copied into that graph's static device
buffers (pinned host -> device, asynchronous) and replayed — ~1 launch instead of ~1900 — and an unseen layout runs one eager
iteration (which also builds its plans and packed operands) and is captured for the next time it appears.

The captured iteration contains everything `TrainStep.step` does: attribute estimation, D-step, the three discriminator
Adam updates, the in-place re-pack of their GEMM operands, G-step, generator Adam + re-pack, and — under data parallelism — the
bucketed gradient all-reduces.  The CropEncoder noise must come from a device generator (`eps_source`) for the replays to
draw fresh noise; with the reference's CPU-RNG noise an iteration cannot be captured and runs eagerly."""
from __future__ import annotations

import gc
from collections import OrderedDict
from typing import Dict, Optional, Tuple

import torch

from . import ops


def layout_key(batch: Dict[str, torch.Tensor]) -> Tuple:
    o2i = batch["obj_to_img"]
    n = int(batch["imgs"].shape[0])
    counts = torch.bincount(o2i.cpu() if o2i.is_cuda else o2i, minlength=n).tolist()
    return (n, tuple(int(c) for c in counts), tuple(batch["imgs"].shape[1:]))


class CapturedIteration:
    """one layout: pinned host staging buffers, static device batch, captured graph (or None: eager)"""

    def __init__(self, ts, host: Dict[str, torch.Tensor], capture: bool, thread_local: bool, warm: int = 1,
                 side_warm: bool = True):
        self.ts = ts
        self.pinned = {k: (v if k == "obj_to_img" else v.contiguous().pin_memory()) for k, v in host.items()}
        self.h2d_bytes = sum(v.numel() * v.element_size() for k, v in host.items() if k != "obj_to_img")
        self.n_images, self.n_objs = int(host["imgs"].shape[0]), int(host["objs"].shape[0])
        self.b = ts.to_device(self.pinned)
        self.graph, self.out, self.error, self.eager_result = None, None, None, None
        self.fork_error = None
        for _ in range(warm):
            self.eager_result = ts.step(self.b, optimizer_step=True)
        torch.cuda.synchronize()
        if not capture:
            return
        gc.collect()
        try:
            if side_warm:            # one more iteration on a side stream (PyTorch's capture recipe); optional for this library,
                side = torch.cuda.Stream()       # whose kernels keep no per-stream state
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    ts.step(self.b, optimizer_step=True)
                torch.cuda.current_stream().wait_stream(side)
                torch.cuda.synchronize()
            self._capture(thread_local)
        except Exception as e:       # capture is an optimisation, never a correctness requirement
            self.error = "%s: %s" % (type(e).__name__, e)
            self.graph = None
            torch.cuda.synchronize()
            if ops.FORKS != "off":
                # second attempt as a single-stream graph (without the fork / join branches of DESIGN.md §3)
                prev, ops.FORKS = ops.FORKS, "off"
                try:
                    self._capture(thread_local)
                    self.fork_error, self.error = self.error, None       # captured, but as a single-stream graph
                except Exception as e2:
                    self.error += " | single-stream retry: %s: %s" % (type(e2).__name__, e2)
                    self.graph = None
                    torch.cuda.synchronize()
                finally:
                    ops.FORKS = prev

    def _capture(self, thread_local: bool):
        g = torch.cuda.CUDAGraph()
        # thread_local: the NCCL watchdog thread may query events while a step with all-reduces is being captured
        with torch.cuda.graph(g, capture_error_mode="thread_local" if thread_local else "global"):
            r = self.ts.step(self.b, optimizer_step=True)
            self.out = dict(d_loss=r["d_loss"], g_loss=r["g_loss"], d_terms=r["d_terms"], g_terms=r["g_terms"],
                            out_g=r["out_g"])
        self.graph = g

    def upload(self, host: Optional[Dict[str, torch.Tensor]] = None):
        """pinned host batch -> static device batch (asynchronous); `host`: new values for the same layout"""
        if host is not None:
            for k, v in host.items():
                if k != "obj_to_img":
                    self.pinned[k].copy_(v)
        for k, v in self.pinned.items():
            if k != "obj_to_img":
                self.b[k].copy_(v, non_blocking=True)

    def run(self):
        if self.graph is not None:
            self.graph.replay()
            return self.out
        return self.ts.step(self.b, optimizer_step=True)


class GraphedTrainStep:
    def __init__(self, ts, max_layouts: int = 8, capture: bool = True):
        self.ts, self.max_layouts, self.capture = ts, max_layouts, capture
        self.cache: "OrderedDict[Tuple, CapturedIteration]" = OrderedDict()
        self.hits = self.misses = 0

    def step(self, host_batch: Dict[str, torch.Tensor]):
        """host_batch: CPU tensors as the loader yields them (train64.py:141 + z).  Returns the dict TrainStep.step returns
        (static tensors of the layout's graph: copy what must outlive the next call)."""
        key = layout_key(host_batch)
        it = self.cache.get(key)
        if it is None:
            self.misses += 1
            if len(self.cache) >= self.max_layouts:
                self.cache.popitem(last=False)
                gc.collect()
            # the eager iteration inside the constructor IS this batch's training iteration; the capture that follows records
            # the launches without executing them, so an unseen layout costs exactly one (eager) iteration
            it = self.cache[key] = CapturedIteration(self.ts, host_batch, self.capture, self.ts.ddp_d is not None, warm=1,
                                                     side_warm=False)
            return it.eager_result
        self.hits += 1
        self.cache.move_to_end(key)
        it.upload(host_batch)
        return it.run()
