"""Layer modules: subclasses of the torch.nn layers the reference uses (same constructors, default init and
state_dict keys) whose forward runs the libb200gan kernels on channel-last activations."""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .ops import ConvGeom, WeightPacks


def _sn_call(m, groups=1):
    """The spectral-norm evaluation this forward uses: the one `sn_prepare` staged for the layer, else `groups` power
    iterations run here (torch.nn.utils.spectral_norm's pre-forward hook, once per batched call)."""
    if not getattr(m, "_b200_sn", False):
        return None
    staged = m.__dict__.pop("_sn_staged", None)
    if staged is not None and staged.groups == groups:
        return staged
    return ops.sn_iterate(m.weight_orig, m.weight_u, m.weight_v, groups, m.training)


def sn_prepare(net, groups=1):
    """Run the power iterations of every spectral-normalised layer below `net` up front (they depend on the weights
    only), `groups` per layer for `groups` batched calls, and stage the results for the layers' next forward."""
    layers = net.__dict__.get("_sn_layers")
    if layers is None:
        layers = [m for m in net.modules() if getattr(m, "_b200_sn", False)]
        net.__dict__["_sn_layers"] = layers
        net.__dict__["_sn_plans"] = {}
    if not layers:
        return
    plans = net.__dict__["_sn_plans"]
    plan = plans.get(groups)
    if plan is None or not plan.valid_for(layers):
        plan = plans[groups] = ops.SNPlan(layers, groups)
    for m, call in zip(layers, plan.run(net.training)):
        m.__dict__["_sn_staged"] = call


def _weight(m):
    return m.weight_orig if getattr(m, "_b200_sn", False) else m.weight


class Conv2d(nn.Conv2d):
    """nn.Conv2d with square kernel / stride / padding as used on the path."""

    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        self._geom = ConvGeom(self.in_channels, self.out_channels, self.kernel_size[0], self.kernel_size[1],
                              self.stride[0], self.padding[0])
        self._packs = WeightPacks()
        self._geom_pool = None

    def forward(self, x, x_layout="cl", out_layout="cl", relu=False, groups=1, out_dtype=None, mask_input_grad=False,
                grad_premasked=False, stats=False, pool=False):
        """groups: number of independent calls batched along dim 0 (each gets its own spectral-norm iteration);
        out_dtype: storage type of a channel-last output (default: that of a channel-last input, else ops.act_dtype());
        relu + grad_premasked on a producer and mask_input_grad on its ONLY consumer fuse the ReLU backward into the
        consumer's data-gradient GEMM (see ops._ConvFn); stats: the output goes straight into a batch norm — let the
        convolution's epilogue accumulate its statistics (ops.conv2d); pool: return avg_pool2d(conv(x), 2) — computed as
        ONE stride-2 convolution with the folded weight (ops.ConvGeom.pooled), the full-resolution output never exists"""
        geom = self._geom
        if pool:
            assert self.stride[0] == 1 and not relu and not stats, "pool=True: stride-1 convolution without fused ReLU"
            if self._geom_pool is None:
                self._geom_pool = ConvGeom.pooled(self.in_channels, self.out_channels, self.kernel_size[0],
                                                  self.kernel_size[1], self.padding[0])
            geom = self._geom_pool
        return ops.conv2d(x, _weight(self), self.bias, geom, self._packs, x_layout, out_layout, relu,
                          _sn_call(self, groups), out_dtype, mask_input_grad, grad_premasked, stats)


class ConvTranspose2d(nn.ConvTranspose2d):
    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        assert self.bias is None, "ConvTranspose2d on the path has no bias"
        # conv orientation: Y = this layer's input (in_channels), X = its output (out_channels)
        self._geom = ConvGeom(self.out_channels, self.in_channels, self.kernel_size[0], self.kernel_size[1],
                              self.stride[0], self.padding[0])
        self._packs = WeightPacks()

    def forward(self, x):
        N, H, W, _ = x.shape
        s, p, k = self.stride[0], self.padding[0], self.kernel_size[0]
        out_hw = ((H - 1) * s - 2 * p + k, (W - 1) * s - 2 * p + k)
        return ops.conv_transpose2d(x, self.weight, self._geom, self._packs, out_hw)


class Linear(nn.Linear):
    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        self._packs = WeightPacks()

    def forward(self, x, relu=False, groups=1, out_dtype=None):
        return ops.linear(x, _weight(self), self.bias, self._packs, relu, _sn_call(self, groups), None, out_dtype)


# ---- num_batches_tracked bookkeeping -----------------------------------------------------------------------------
# Every training-mode BN forward increments its int64 counter (nn.BatchNorm semantics).  Inside a training step the ~70
# one-element increments are deferred and applied with ONE multi-tensor add when the step ends (count_batches / flush).
_DEFER_COUNTS = False
_PENDING_COUNTS = {}


def count_batches(counter: torch.Tensor, n: int):
    if _DEFER_COUNTS:
        key = id(counter)
        ent = _PENDING_COUNTS.get(key)
        if ent is None:
            _PENDING_COUNTS[key] = [counter, n]
        else:
            ent[1] += n
    else:
        counter.add_(n)


class deferred_batch_counts:
    """with deferred_batch_counts(): ...  — the counters are brought up to date on exit"""

    def __enter__(self):
        global _DEFER_COUNTS
        self.prev, _DEFER_COUNTS = _DEFER_COUNTS, True
        return self

    def __exit__(self, *exc):
        global _DEFER_COUNTS
        _DEFER_COUNTS = self.prev
        if not _DEFER_COUNTS and _PENDING_COUNTS:
            ents = list(_PENDING_COUNTS.values())
            _PENDING_COUNTS.clear()
            torch._foreach_add_([e[0] for e in ents], [e[1] for e in ents])
        return False


class BatchNorm2d(nn.BatchNorm2d):
    """Works on (..., C) channel-last tensors (also serves BatchNorm1d's (B, C) case)."""

    def forward(self, x, relu=False, residual=None, groups=1):
        if self.training and self.track_running_stats:
            count_batches(self.num_batches_tracked, groups)
        w = self.weight if self.affine else None
        b = self.bias if self.affine else None
        return ops.batch_norm(x, w, b, self.running_mean, self.running_var, self.training, relu, residual, groups)


class BatchNorm1d(nn.BatchNorm1d):
    def forward(self, x, relu=False, groups=1):
        if self.training and self.track_running_stats:
            count_batches(self.num_batches_tracked, groups)
        return ops.batch_norm(x, self.weight, self.bias, self.running_mean, self.running_var, self.training, relu, None,
                              groups)


class Embedding(nn.Embedding):
    def forward(self, idx_i32):
        return ops.embedding(self.weight, idx_i32)


class ReLU(nn.ReLU):
    def forward(self, x):
        return ops.relu(x)


def add_sn(m):
    """discriminator.py:15-22 — wrap every Conv2d / ConvTranspose2d / Linear / Embedding below `m` with spectral
    normalisation.  State layout equals torch.nn.utils.spectral_norm's (weight_orig parameter, weight_u / weight_v
    buffers, u ~ normalize(N(0,1)) of size Cout, v of size Cin*kh*kw); the power iteration, the 1/sigma scaling and
    the gradient through sigma run in libb200gan (b200_sn_power_iter, conv epilogue scale, b200_sn_grad)."""
    for name, c in m.named_children():
        m.add_module(name, add_sn(c))
    if isinstance(m, (Conv2d, Linear)):
        if getattr(m, "_b200_sn", False):
            raise RuntimeError("spectral norm already applied")
        w = m._parameters.pop("weight")
        m.register_parameter("weight_orig", w)
        h, wd = w.shape[0], w[0].numel()
        u = F.normalize(w.new_empty(h).normal_(0, 1), dim=0, eps=ops.SN_EPS)
        v = F.normalize(w.new_empty(wd).normal_(0, 1), dim=0, eps=ops.SN_EPS)
        m.register_buffer("weight_u", u)
        m.register_buffer("weight_v", v)
        m._b200_sn = True
        return m
    if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d, nn.Linear, nn.Embedding)):
        raise NotImplementedError("add_sn: %s has no libb200gan spectral-norm path" % type(m).__name__)
    return m
