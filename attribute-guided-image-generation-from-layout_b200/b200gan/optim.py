"""Multi-tensor Adam on the libb200gan kernel: the four torch.optim.Adam instances of
train64.py:111-114 (lr, betas=(0.5, 0.999), eps=1e-8, no weight decay, no amsgrad) as ONE launch per optimizer whose step
counter lives on the device, so the whole update is CUDA-graph capturable.

The class keeps torch.optim.Optimizer's surface (param_groups, state[p] = {"step", "exp_avg", "exp_avg_sq"}, zero_grad,
state_dict / load_state_dict — the latter rebuilds the device step counter and tables), so utils/model_saver_iter.py
style checkpointing works unchanged.  All parameters of one
instance share one step counter (every parameter of a network receives a gradient in every step of the reference's loop)."""
import ctypes as C

import torch

from . import _lib

CHUNK = 65536


class _Entry(C.Structure):
    _fields_ = [("param", C.c_void_p), ("grad", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p),
                ("n", C.c_int32), ("pad", C.c_int32)]


class Adam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))
        self._tables = {}
        self._captured = {}
        self._steps = {}

    def _table(self, group_idx, items):
        """device array of b200_adam_entry for the group's chunks, keyed by the tensors' addresses.

        Eager iterations share ONE reusable slot per group (pinned host image + device buffer; gradient tensors move from
        step to step, so the image is rewritten and re-copied with an asynchronous pinned copy).  A table built WHILE A CUDA
        GRAPH IS BEING CAPTURED gets its own pinned image and device buffer that are never rewritten: the captured memcpy
        node re-reads the host image on every replay, so sharing it between captures (several batch layouts, each with its own
        graph and its own gradient buffers) would make one graph's Adam read another graph's gradients."""
        key = tuple((p.data_ptr(), p.grad.data_ptr(), self.state[p]["exp_avg"].data_ptr(),
                     self.state[p]["exp_avg_sq"].data_ptr()) for p in items)
        capturing = torch.cuda.is_current_stream_capturing()
        n_ent = sum(-(-p.numel() // CHUNK) for p in items)
        nbytes = n_ent * C.sizeof(_Entry)
        if capturing:
            slot = self._captured.get((group_idx, key))
            if slot is not None:
                return slot["dev"], slot["n"]
            slot = dict(key=None, n=0, host=torch.empty((nbytes,), dtype=torch.uint8).pin_memory(),
                        dev=torch.empty((nbytes,), dtype=torch.uint8, device=items[0].device))
            self._captured[(group_idx, key)] = slot
        else:
            slot = self._tables.get(group_idx)
            if slot is None or slot["host"].numel() < nbytes:
                slot = dict(key=None, n=0, host=torch.empty((nbytes,), dtype=torch.uint8).pin_memory(),
                            dev=torch.empty((nbytes,), dtype=torch.uint8, device=items[0].device))
                self._tables[group_idx] = slot
        if slot["key"] != key:
            if not capturing and slot.get("evt") is not None:
                slot["evt"].synchronize()          # the previous copy of this image must have left the host buffer (an event
                                                   # wait on that copy alone: a stream synchronize here stalled the host per step)
            arr = (_Entry * n_ent).from_address(slot["host"].data_ptr())
            i = 0
            for p in items:
                st = self.state[p]
                n = p.numel()
                for off in range(0, n, CHUNK):
                    arr[i] = _Entry(p.data_ptr() + off * 4, p.grad.data_ptr() + off * 4, st["exp_avg"].data_ptr() + off * 4,
                                    st["exp_avg_sq"].data_ptr() + off * 4, min(CHUNK, n - off), 0)
                    i += 1
            slot["dev"][:nbytes].copy_(slot["host"][:nbytes], non_blocking=True)
            if not capturing:
                slot["evt"] = torch.cuda.Event()
                slot["evt"].record()
            slot["key"], slot["n"] = key, n_ent
        return slot["dev"], slot["n"]

    def load_state_dict(self, state_dict):
        """torch.optim.Optimizer.load_state_dict, then: the kernel reads ONE device step counter per group, so it is
        rebuilt from the loaded per-parameter `step` values (bias correction continues where the checkpoint stopped), and
        the cached device tables are dropped (the loaded exp_avg / exp_avg_sq are new tensors)."""
        super().load_state_dict(state_dict)
        self._tables = {}
        self._captured = {}
        self._steps = {}
        for gi, group in enumerate(self.param_groups):
            ps = [p for p in group["params"] if p in self.state and "step" in self.state[p]]
            if not ps:
                continue
            t = max(float(self.state[p]["step"]) for p in ps)
            step_t = torch.full((), t, dtype=torch.float32, device=ps[0].device)
            self._steps[gi] = step_t
            for p in ps:
                st = self.state[p]
                st["step"] = step_t
                for k in ("exp_avg", "exp_avg_sq"):
                    if st[k].device != p.device or st[k].dtype != torch.float32 or not st[k].is_contiguous():
                        st[k] = st[k].to(device=p.device, dtype=torch.float32).contiguous()

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for gi, group in enumerate(self.param_groups):
            items = [p for p in group["params"] if p.grad is not None]
            if not items:
                continue
            step_t = self._steps.get(gi)
            if step_t is None:
                step_t = self._steps[gi] = torch.zeros((), dtype=torch.float32, device=items[0].device)
            for p in items:
                if p.dtype != torch.float32 or not p.is_cuda or not p.is_contiguous() or not p.grad.is_contiguous():
                    raise _lib.B200Error("b200gan.optim.Adam: contiguous fp32 CUDA parameters and gradients required")
                st = self.state[p]
                if not st:
                    st["step"] = step_t
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            table, n = self._table(gi, items)
            b1, b2 = group["betas"]
            _lib.K.adam_multi(table, n, step_t, group["lr"], b1, b2, group["eps"])
        return loss
