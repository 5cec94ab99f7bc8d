"""torch.autograd.Function layer over the libb200gan C ABI.

Tensor conventions: "cl" (channel-last) activations are contiguous (N, H, W, C) fp32 tensors; "nchw" tensors are
the reference-boundary layout (3-channel images / crops) and are addressed through strides, never transposed.
Every forward and backward here launches only kernels of this repository (plus torch allocations / views).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import contextlib
import os
import weakref

import torch

from . import _lib
from ._lib import ConvDesc, MODE_AFFINE, MODE_CBN, MODE_PLAIN, MODE_SPADE

BN_EPS = 1e-5
BN_MOMENTUM = 0.1
SN_EPS = 1e-12

# ----------------------------------------------------------------------------------------------------------
# precision policy: "fp32" = CUDA-core fp32 gather-GEMMs (tight parity); "bf16" = tcgen05 bf16 operands with
# fp32 accumulation wherever the layer shape is eligible (channels % 64 == 0, channel-last), fp32 otherwise;
# "tf32" = fp32 tensors in memory, tcgen05 kind::tf32 GEMMs (operands rounded to a 10-bit mantissa on the way into shared
# memory, fp32 accumulation) where channels % 32 == 0 — BASELINE config 2's "fp32" half on the tensor cores.
# ----------------------------------------------------------------------------------------------------------
_PRECISION = "fp32"
_CACHE_EPOCH = 0
_WGRAD_TARGET = int(os.environ.get("B200_WGRAD_TARGET", "296"))     # experiments only (read once: os.environ is slow)


def set_precision(p: str):
    global _PRECISION
    if p not in ("fp32", "bf16", "tf32"):
        raise ValueError("precision must be 'fp32', 'tf32' or 'bf16'")
    _PRECISION = p


def get_precision() -> str:
    return _PRECISION


def act_dtype() -> torch.dtype:
    """storage type of channel-last activations (and their gradients) created by this package: bf16 in the bf16 mode,
    fp32 otherwise.  NCHW boundary tensors, statistics, parameters and parameter gradients are always fp32."""
    return torch.bfloat16 if _PRECISION == "bf16" else torch.float32


def bump_weight_epoch():
    """Invalidate every packed-weight cache entry (call after weights change behind autograd's back, e.g. a CUDA
    graph replay that contains the optimizer step)."""
    global _CACHE_EPOCH
    _CACHE_EPOCH += 1


def _rup(x: int, m: int) -> int:
    return (x + m - 1) // m * m


def cl_strides(H: int, W: int, C: int):
    return (H * W * C, W * C, C, 1)


def nchw_strides(C: int, H: int, W: int):
    return (C * H * W, W, 1, H * W)  # (sn, sh, sw, sc)


def _dims(x: torch.Tensor, layout: str):
    if layout == "cl":
        N, H, W, C = x.shape
        return N, H, W, C, cl_strides(H, W, C)
    N, C, H, W = x.shape
    return N, H, W, C, nchw_strides(C, H, W)


def _empty(N, H, W, C, layout, device, dtype=torch.float32):
    if layout == "cl":
        return torch.empty((N, H, W, C), dtype=dtype, device=device), cl_strides(H, W, C)
    return torch.empty((N, C, H, W), dtype=torch.float32, device=device), nchw_strides(C, H, W)


def _out_dtype(x: torch.Tensor, x_layout: str, out_dtype: Optional[torch.dtype]) -> torch.dtype:
    """channel-last outputs keep the storage type of a channel-last input; an NCHW (boundary) input starts a network,
    whose activations take act_dtype()"""
    if out_dtype is not None:
        return out_dtype
    return x.dtype if x_layout == "cl" else act_dtype()


class ConvGeom:
    """A convolution in *conv orientation*: X (N,Hx,Wx,Cx) * W (Cy,Cx,kh,kw) -> Y (N,Hy,Wy,Cy).
    nn.ConvTranspose2d is the same geometry run backwards (its input is Y, its output X)."""

    def __init__(self, Cx, Cy, kh, kw, stride, pad, cx_offset=0, cx_total=None, fold=False):
        self.Cx, self.Cy, self.kh, self.kw, self.s, self.p = Cx, Cy, kh, kw, stride, pad
        # the parameter may hold more input channels than this op uses (ConvLSTM x / h halves)
        self.cx_offset = cx_offset
        self.cx_total = Cx if cx_total is None else cx_total
        # fold: this is the stride-2 form of avg_pool2(conv_{(kh-1) x (kw-1), stride 1}) — the PARAMETER has (kh-1) x (kw-1)
        # taps and the GEMM operands are packed from its folded image (ConvGeom.pooled, b200_fold_pool_weight)
        self.fold = bool(fold)

    @classmethod
    def pooled(cls, Cx, Cy, kh, kw, pad):
        """geometry of avg_pool2(conv_{kh x kw, stride 1, pad}(x)) as ONE (kh+1) x (kw+1) stride-2 convolution"""
        return cls(Cx, Cy, kh + 1, kw + 1, 2, pad, fold=True)

    def out_hw(self, Hx, Wx):
        return (Hx + 2 * self.p - self.kh) // self.s + 1, (Wx + 2 * self.p - self.kw) // self.s + 1

    def key(self):
        return (self.Cx, self.Cy, self.kh, self.kw, self.s, self.p, self.cx_offset, self.cx_total, self.fold)


class PackRecipe:
    """One b200_pack_weight call: how a GEMM operand matrix (or a row / channel slice of it) is produced from a parameter.
    Kept with the cached matrix so that the whole set can be re-packed in place by ONE launch after an optimizer step
    (refresh_packs -> b200_pack_weight_multi)."""
    __slots__ = ("src", "src_offset", "dst", "dst_row_offset", "bf16", "M", "Mpad", "Th", "Tw", "C", "ldw", "s_m", "s_ky",
                 "s_kx", "s_c", "ky0", "kx0", "kstep", "C_dst", "c_off")

    def __init__(self, src, src_offset, dst, dst_row_offset, bf16, M, Mpad, Th, Tw, C, ldw, s_m, s_ky, s_kx, s_c, ky0=0,
                 kx0=0, kstep=1, C_dst=0, c_off=0):
        (self.src, self.src_offset, self.dst, self.dst_row_offset, self.bf16, self.M, self.Mpad, self.Th, self.Tw, self.C,
         self.ldw, self.s_m, self.s_ky, self.s_kx, self.s_c, self.ky0, self.kx0, self.kstep, self.C_dst, self.c_off) = (
            src, src_offset, dst, dst_row_offset, bool(bf16), M, Mpad, Th, Tw, C, ldw, s_m, s_ky, s_kx, s_c, ky0, kx0,
            kstep, C_dst, c_off)

    def run(self):
        _lib.K.pack_weight(self.src, self.src_offset, self.dst, self.dst_row_offset, self.bf16, self.M, self.Mpad, self.Th,
                           self.Tw, self.C, self.ldw, self.s_m, self.s_ky, self.s_kx, self.s_c, self.ky0, self.kx0,
                           self.kstep, self.C_dst, self.c_off)


class FoldRecipe:
    """w4 = fold(w): the weight of the stride-2 convolution equal to avg_pool2(conv(x; w)); runs BEFORE the PackRecipes that
    read w4 (WeightPacks.get keeps it first in the entry's list; refresh_packs launches the folds, then the packing)"""
    __slots__ = ("src", "dst", "__weakref__")

    def __init__(self, src, dst):
        self.src, self.dst = src, dst

    def run(self):
        _lib.K.fold_pool_weight(self.src, self.dst)


_FOLD_IMAGES = weakref.WeakValueDictionary()      # parameter storage -> its live FoldRecipe (shared by the layer's operands)


def _folded_source(g: ConvGeom, srcs, recipes):
    """the (Cy, cx_total, kh, kw) fp32 image of a pooled convolution's parameter that the packing reads; the forward and the
    data-gradient operands of a layer share one image (one fold per weight update)"""
    assert len(srcs) == 1 and g.fold
    w = srcs[0]
    assert tuple(w.shape[2:]) == (g.kh - 1, g.kw - 1) and w.is_contiguous(), (tuple(w.shape), g.key())
    key = (w.data_ptr(), tuple(w.shape), str(w.device))
    fr = _FOLD_IMAGES.get(key)
    if fr is None:
        w4 = torch.empty((w.shape[0], w.shape[1], g.kh, g.kw), dtype=torch.float32, device=w.device)
        fr = _FOLD_IMAGES[key] = FoldRecipe(w, w4)
    fr.run()
    recipes.append(fr)
    return (fr.dst,)


def unfold_pool_grad(g4: torch.Tensor) -> torch.Tensor:
    """transpose of the fold: the gradient of the (kh, kw) parameter from that of its folded (kh+1, kw+1) image"""
    return 0.25 * ((g4[..., :-1, :-1] + g4[..., :-1, 1:]) + (g4[..., 1:, :-1] + g4[..., 1:, 1:]))


# Discriminator blocks evaluate avg_pool2(conv(h)) as one stride-2 convolution with the folded weight and the 1x1 shortcut on
# the pooled input (models/discriminator.py).  B200_POOLED_CONV=0 keeps the literal operation order.
POOLED_CONV = os.environ.get("B200_POOLED_CONV", "1") != "0"

_PACK_REGISTRY = weakref.WeakSet()      # every live WeightPacks
_PACK_GEN = 0                            # bumped whenever a cached operand is (re)built into a new buffer
_REFRESH_TABLES: Dict[object, tuple] = {}


class WeightPacks:
    """Packed GEMM operands of one layer's weights.

    Validity: an entry is rebuilt lazily when its source parameters' (data_ptr, _version) or the global epoch changed
    (in-place autograd-visible updates: load_state_dict, foreach/single-tensor optimizers, init code), and ALL entries
    derived from an optimizer's parameters are re-packed eagerly, in place, by one launch right after every
    Optimizer.step() (refresh_packs, installed as a global optimizer post-step hook by the package) — which covers
    updates that do not touch `_version` (b200gan.optim.Adam's raw-pointer kernel, torch's fused Adam, a CUDA-graph
    replay containing the update).  Anything else that writes weights behind autograd's back (`.data`, raw pointers)
    must call bump_weight_epoch()."""

    def __init__(self, sources=None):
        self.store: Dict[tuple, list] = {}
        # parameters whose concatenation along dim 0 is the weight `w` handed to get() (SPADE's fused gamma|beta
        # convolution): the operand is packed from them directly and `w` itself may be a per-call temporary
        self.sources = tuple(sources) if sources is not None else None
        _PACK_REGISTRY.add(self)

    @staticmethod
    def _stamp(src):
        return tuple((t.data_ptr(), t._version) for t in src) + (_CACHE_EPOCH,)

    def get(self, key, w: torch.Tensor, builder):
        """builder(sources, recipes) -> value; appends one PackRecipe per b200_pack_weight call it made"""
        global _PACK_GEN
        src = self.sources if self.sources is not None else (w,)
        stamp = self._stamp(src)
        hit = self.store.get(key)
        if hit is not None and hit[0] == stamp:
            return hit[1]
        if hit is not None and all(a[0] == b[0] for a, b in zip(hit[0][:-1], stamp[:-1])) and len(hit[3]) == len(src):
            for r in hit[2]:                       # same parameter storage, new values: re-pack in place
                r.run()
            hit[0] = stamp
            return hit[1]
        # the cache outlives the iteration: hold graph-free aliases (a view of a parameter made under grad mode would keep
        # its AccumulateGrad node — and its stream — alive across iterations, which breaks CUDA-graph capture)
        src = tuple(t.detach() for t in src)
        recipes: List[PackRecipe] = []
        val = builder(src, recipes)
        self.store[key] = [stamp, val, recipes, src]
        _PACK_GEN += 1
        return val


def _storage_ptr(t: torch.Tensor) -> int:
    return t.untyped_storage().data_ptr()


def refresh_packs(params=None, owner=None):
    """Re-pack, in place and with ONE launch, every cached GEMM operand derived from `params` (all cached operands when
    None) and mark them valid for the parameters' current versions.  Called after every optimizer step (see
    WeightPacks); CUDA-graph capturable (the device table is cached per `owner` — the optimizer — and operand-set
    generation, so no host-to-device copy happens in steady state)."""
    ent = _REFRESH_TABLES.get(id(owner)) if owner is not None else None
    if ent is None or ent[0] != _PACK_GEN or ent[3]() is not owner:
        ptrs = None if params is None else {_storage_ptr(p) for p in params}
        items = []
        for wp in list(_PACK_REGISTRY):
            for e in wp.store.values():
                if ptrs is None or any(_storage_ptr(t) in ptrs for t in e[3]):
                    items.append(e)
        every = [r for e in items for r in e[2]]
        folds = list({id(r): r for r in every if isinstance(r, FoldRecipe)}.values())
        recipes = [r for r in every if not isinstance(r, FoldRecipe)]
        table = _lib.K.pack_table(recipes, recipes[0].src.device) if recipes else None
        ent = (_PACK_GEN, items, table, weakref.ref(owner) if owner is not None else None, folds)
        if owner is not None:
            if len(_REFRESH_TABLES) > 64:
                _REFRESH_TABLES.clear()
            _REFRESH_TABLES[id(owner)] = ent
    _, items, table, _, folds = ent
    if table is None:
        return 0
    for fr in folds:                       # pooled convolutions: the folded weight images first, the packing reads them
        fr.run()
    _host, dev, n, chunks = table
    _lib.K.pack_weight_multi(dev, n, chunks)
    for e in items:
        e[0] = WeightPacks._stamp(e[3])
    return n


def _optimizer_post_step(optimizer, args, kwargs):
    """global torch.optim post-step hook: the packed operands follow every parameter update (see WeightPacks)"""
    if not any(len(wp.store) for wp in _PACK_REGISTRY):
        return
    refresh_packs([p for g in optimizer.param_groups for p in g["params"]], owner=optimizer)


_HOOK_HANDLE = None


def install_optimizer_hook():
    global _HOOK_HANDLE
    if _HOOK_HANDLE is None:
        from torch.optim.optimizer import register_optimizer_step_post_hook
        _HOOK_HANDLE = register_optimizer_step_post_hook(_optimizer_post_step)


TC_NONE, TC_BF16, TC_TF32 = 0, 1, 2     # GEMM kernel family of a launch: CUDA-core fp32 | tcgen05 bf16 | tcgen05 tf32


def _tc_kind() -> int:
    return TC_BF16 if _PRECISION == "bf16" else (TC_TF32 if _PRECISION == "tf32" else TC_NONE)


def _kalign(kind: int) -> int:
    """channels per 128-byte operand row: the K granularity of the tcgen05 kernels"""
    return 64 if kind == TC_BF16 else 32


def tc_operand(x: torch.Tensor, kind: int) -> torch.Tensor:
    """the activation operand a tcgen05 GEMM of `kind` consumes: a bf16 copy (TC_BF16) or the fp32 tensor itself"""
    return as_bf16(x) if kind == TC_BF16 else x


def _tc_fwd_ok(g: ConvGeom, x_layout: str) -> int:
    k = _tc_kind()
    return k if (k and x_layout == "cl" and g.Cx % _kalign(k) == 0) else TC_NONE


def _tc_dgrad_ok(g: ConvGeom, dy_layout: str) -> int:
    k = _tc_kind()
    return k if (k and dy_layout == "cl" and g.Cy % _kalign(k) == 0) else TC_NONE


def _tc_wgrad_ok(g: ConvGeom, x_layout: str, dy_layout: str) -> int:
    """the weight gradient needs pixel-major (MN-major) operands; in the tf32 mode it runs on the bf16 kernel as three GEMMs
    over two-term bf16 splits of the fp32 operands (conv_wgrad), hence the 64-channel granularity in both modes"""
    k = _tc_kind()
    return k if (k and x_layout == "cl" and dy_layout == "cl" and g.Cx % 64 == 0 and g.Cy % 64 == 0) else TC_NONE


def _pack_fwd(g: ConvGeom, srcs, tc: bool, recipes: List[PackRecipe]):
    """wmat[co][(ky*kw+kx)*Cx + c] = w[co, cx_offset+c, ky, kx]; w = srcs concatenated along dim 0"""
    if g.fold:
        srcs = _folded_source(g, srcs, recipes)
    K = g.kh * g.kw * g.Cx
    ldw = _rup(K, _kalign(tc)) if tc else _rup(K, 4)
    mpad = _rup(g.Cy, _lib.K.conv_tc_ntile(g.Cy)) if tc else g.Cy
    dst = torch.zeros((mpad, ldw), dtype=torch.bfloat16 if tc == TC_BF16 else torch.float32, device=srcs[0].device)
    kk = g.kh * g.kw
    row = 0
    for i, w in enumerate(srcs):
        rows = w.shape[0]
        last = i == len(srcs) - 1
        r = PackRecipe(w, g.cx_offset * kk, dst, row, tc == TC_BF16, rows, (mpad - row) if last else rows, g.kh, g.kw, g.Cx, ldw,
                       g.cx_total * kk, g.kw, 1, kk)
        r.run()
        recipes.append(r)
        row += rows
    assert row == g.Cy, (row, g.Cy)
    return dst, ldw


def _dgrad_phases(g: ConvGeom):
    out = []
    for py in range(g.s):
        for px in range(g.s):
            ky0, kx0 = (py + g.p) % g.s, (px + g.p) % g.s
            Th = (g.kh - ky0 + g.s - 1) // g.s if g.kh > ky0 else 0
            Tw = (g.kw - kx0 + g.s - 1) // g.s if g.kw > kx0 else 0
            out.append((py, px, ky0, kx0, Th, Tw))
    return out


def _pack_dgrad(g: ConvGeom, srcs, tc: bool, recipes: List[PackRecipe]):
    """per output phase: wmat[ci][(j*Tw+i)*Cy + co] = w[co, cx_offset+ci, ky0+s*j, kx0+s*i]; w = srcs concatenated along
    dim 0 (each source fills its own slice of the co axis)"""
    if g.fold:
        srcs = _folded_source(g, srcs, recipes)
    packs = []
    kk = g.kh * g.kw
    for (py, px, ky0, kx0, Th, Tw) in _dgrad_phases(g):
        if Th == 0 or Tw == 0:
            packs.append(None)
            continue
        K = Th * Tw * g.Cy
        ldw = _rup(K, _kalign(tc)) if tc else _rup(K, 4)
        mpad = _rup(g.Cx, _lib.K.conv_tc_ntile(g.Cx)) if tc else g.Cx
        dst = torch.zeros((mpad, ldw), dtype=torch.bfloat16 if tc == TC_BF16 else torch.float32, device=srcs[0].device)
        co = 0
        for w in srcs:
            rows = w.shape[0]
            whole = len(srcs) == 1
            r = PackRecipe(w, g.cx_offset * kk, dst, 0, tc == TC_BF16, g.Cx, mpad, Th, Tw, rows, ldw, kk, g.kw, 1, g.cx_total * kk,
                           ky0, kx0, g.s, 0 if whole else g.Cy, 0 if whole else co)
            r.run()
            recipes.append(r)
            co += rows
        assert co == g.Cy, (co, g.Cy)
        packs.append((dst, ldw))
    return packs


def as_bf16(x: torch.Tensor) -> torch.Tensor:
    """bf16 operand copy of a channel-last activation for the tcgen05 gather-GEMMs (no-op when already bf16)"""
    if x.dtype == torch.bfloat16:
        return x
    return _lib.K.cast_bf16(x.contiguous())


def _scale_rows(scale, rows: int) -> int:
    """rows of the GEMM that share one entry of `scale` (0 = a single scalar): batched spectral-norm calls"""
    if scale is None or scale.numel() == 1:
        return 0
    assert rows % scale.numel() == 0, (rows, scale.numel())
    return rows // scale.numel()


def _packed_ok(g: ConvGeom) -> bool:
    """few-channel inputs (the 3-channel image / crop convolutions) reach the tcgen05 GEMMs through an explicit bf16
    im2col matrix: K = Cx*kh*kw per output pixel, zero padded to a multiple of 64"""
    return (_PRECISION == "bf16" and g.Cx < 16 and g.Cy % 64 == 0 and g.cx_offset == 0 and g.cx_total == g.Cx
            and _rup(g.Cx * g.kh * g.kw, 64) <= 256)


def im2col_pack(g: ConvGeom, x, x_layout):
    """(N*Hy*Wy, Kp) bf16 im2col matrix of x; column (ky*kw+kx)*Cx + c — the column order of _pack_fwd"""
    N, Hx, Wx, Cx, xs = _dims(x, x_layout)
    assert Cx == g.Cx
    Hy, Wy = g.out_hw(Hx, Wx)
    Kp = _rup(g.Cx * g.kh * g.kw, 64)
    return _lib.K.im2col_pack(x.contiguous(), xs, N, Hx, Wx, Cx, g.kh, g.kw, g.s, g.p, Hy, Wy, Kp)


# Batch-norm statistics accumulated in the producing convolution's epilogue (persistent tcgen05 kernel, b200_conv_desc.col_stats
# -> b200_bn_stats_slabs).  Correct (tests/test_kernels_gpu.py::test_conv_epilogue_statistics_*), but measured SLOWER inside the
# step (39.5 vs 38.6 ms): the 32 shuffles per 16 columns roughly double the epilogue of the short-K layers (+0.75 ms over the GEMM
# launches) and the per-32-row partials make the finishing kernel as expensive as the statistics pass it replaces (0.73 ms).  Kept
# as an experiment behind this switch; off by default.  
FUSE_BN_STATS = False


def conv_forward_packed(g: ConvGeom, packs: WeightPacks, w, P, N, Hy, Wy, out_layout, bias=None, scale=None, relu=False,
                        out_dtype=torch.bfloat16, stats_box=None):
    """Y = epilogue(P @ Wmat^T): the convolution as a 1x1 tcgen05 gather-GEMM over the im2col matrix P"""
    Kp = P.shape[1]
    wmat, ldw = packs.get(("fwd", TC_BF16) + g.key(), w, lambda src, rec: _pack_fwd(g, src, TC_BF16, rec))
    assert ldw == Kp
    y, ys = _empty(N, Hy, Wy, g.Cy, out_layout, P.device, out_dtype)
    xs = cl_strides(Hy, Wy, Kp)
    d = ConvDesc(B=N, Qh=Hy, Qw=Wy, Cin=Kp, Cout=g.Cy, Th=1, Tw=1, in_sy=1, in_sx=1, tap_sy=1, tap_sx=1, tap_oy=0,
                 tap_ox=0, Hi=Hy, Wi=Wy, up_shift=0, in_sn=xs[0], in_sh=xs[1], in_sw=xs[2], in_sc=xs[3], out_sy=1, out_sx=1,
                 out_oy=0, out_ox=0, Ho=Hy, Wo=Wy, out_sn=ys[0], out_sh=ys[1], out_sw=ys[2], out_sc=ys[3], ldw=ldw,
                 relu=int(relu), scale_rows=_scale_rows(scale, N * Hy * Wy))
    st = None
    if stats_box is not None and FUSE_BN_STATS and out_layout == "cl" and _lib.K.conv_stats_ok(d, TC_BF16):
        st = _lib.K.conv_stats_buffer(d, P.device)
        stats_box.append(st)
    _lib.K.conv_gemm(d, P, wmat, bias, scale, y, TC_BF16, stats=st)
    return y


def conv_wgrad_packed(g: ConvGeom, P, N, Hy, Wy, dy, dy_layout, dw: torch.Tensor):
    """dW = dY^T @ P on the tcgen05 weight-gradient kernel (T = 1, Kp "channels"), unpacked into the parameter layout"""
    Kp = P.shape[1]
    _, _, _, Cy, ds = _dims(dy, dy_layout)
    assert Cy == g.Cy and dy.dtype == torch.bfloat16
    g1 = ConvGeom(Kp, g.Cy, 1, 1, 1, 0)
    Q = N * Hy * Wy
    splits = _wgrad_splits(g1, Q, TC_BF16)
    ws = torch.empty((splits * g.Cy * Kp,), dtype=torch.float32, device=P.device)
    xs = cl_strides(Hy, Wy, Kp)
    d = ConvDesc(B=N, Qh=Hy, Qw=Wy, Cin=Kp, Cout=g.Cy, Th=1, Tw=1, in_sy=1, in_sx=1, tap_sy=1, tap_sx=1, tap_oy=0,
                 tap_ox=0, Hi=Hy, Wi=Wy, up_shift=0, in_sn=xs[0], in_sh=xs[1], in_sw=xs[2], in_sc=xs[3], out_sy=1, out_sx=1,
                 out_oy=0, out_ox=0, Ho=Hy, Wo=Wy, out_sn=ds[0], out_sh=ds[1], out_sw=ds[2], out_sc=ds[3], ldw=Kp, relu=0)
    _lib.K.wgrad_gemm(d, dy, P, ws, splits, TC_BF16)
    tmp = torch.empty((g.Cy, Kp), dtype=torch.float32, device=P.device)
    _lib.K.wgrad_reduce(ws, splits, g.Cy, 1, 1, Kp, tmp, 0, Kp, 0, 0, 1)
    K = g.kh * g.kw * g.Cx
    dw.copy_(tmp[:, :K].view(g.Cy, g.kh, g.kw, g.Cx).permute(0, 3, 1, 2))
    return dw


def _packed_out_ok(g: ConvGeom, x_layout: str) -> bool:
    """few OUTPUT channels (the 64 -> 3 image convolution): its data and weight gradients share one bf16 im2col matrix of
    the output gradient (transposed window walk) and run on the tcgen05 GEMMs"""
    return (_PRECISION == "bf16" and x_layout == "cl" and g.Cy < 16 and g.Cx % 64 == 0 and g.s == 1 and g.cx_offset == 0
            and g.cx_total == g.Cx and _rup(g.Cy * g.kh * g.kw, 64) <= 256)


def conv_backward_packed_out(g: ConvGeom, packs: WeightPacks, w, x, x_dims, dy, dy_layout, need_dx, need_dw, scale,
                             x_dtype):
    """(dX, dW) of a stride-1 convolution with few output channels through P = im2col(dY) (flipped window):
    dX = P @ Wd^T (a 1x1 tcgen05 gather-GEMM, Wd = the data-gradient weight pack) and dW = X^T @ P (tcgen05 weight-gradient
    kernel with X as the pixel-major M operand and P as the gathered operand, T = 1)."""
    N, Hx, Wx, Cx = x_dims
    _, Hy, Wy, Cy, ds = _dims(dy, dy_layout)
    K = g.kh * g.kw * Cy
    Kp = _rup(K, 64)
    P = _lib.K.im2col_pack(dy, ds, N, Hy, Wy, Cy, g.kh, g.kw, 1, g.p, Hx, Wx, Kp, flip=True)
    ps = cl_strides(Hx, Wx, Kp)
    dx = dw = None
    if need_dx:
        wmat, ldw = packs.get(("dgrad", TC_BF16) + g.key(), w, lambda src, rec: _pack_dgrad(g, src, TC_BF16, rec))[0]
        assert ldw == Kp
        dx, xs = _empty(N, Hx, Wx, Cx, "cl", dy.device, x_dtype)
        d = ConvDesc(B=N, Qh=Hx, Qw=Wx, Cin=Kp, Cout=Cx, Th=1, Tw=1, in_sy=1, in_sx=1, tap_sy=1, tap_sx=1, tap_oy=0,
                     tap_ox=0, Hi=Hx, Wi=Wx, up_shift=0, in_sn=ps[0], in_sh=ps[1], in_sw=ps[2], in_sc=ps[3], out_sy=1,
                     out_sx=1, out_oy=0, out_ox=0, Ho=Hx, Wo=Wx, out_sn=xs[0], out_sh=xs[1], out_sw=xs[2], out_sc=xs[3],
                     ldw=ldw, relu=0, scale_rows=_scale_rows(scale, N * Hx * Wx))
        _lib.K.conv_gemm(d, P, wmat, None, scale, dx, TC_BF16)
    if need_dw:
        xb = as_bf16(x)
        xs = cl_strides(Hx, Wx, Cx)
        g1 = ConvGeom(Kp, Cx, 1, 1, 1, 0)
        splits = _wgrad_splits(g1, N * Hx * Wx, TC_BF16)
        ws = torch.empty((splits * Cx * Kp,), dtype=torch.float32, device=dy.device)
        d = ConvDesc(B=N, Qh=Hx, Qw=Wx, Cin=Kp, Cout=Cx, Th=1, Tw=1, in_sy=1, in_sx=1, tap_sy=1, tap_sx=1, tap_oy=0,
                     tap_ox=0, Hi=Hx, Wi=Wx, up_shift=0, in_sn=ps[0], in_sh=ps[1], in_sw=ps[2], in_sc=ps[3], out_sy=1,
                     out_sx=1, out_oy=0, out_ox=0, Ho=Hx, Wo=Wx, out_sn=xs[0], out_sh=xs[1], out_sw=xs[2], out_sc=xs[3],
                     ldw=Kp, relu=0)
        _lib.K.wgrad_gemm(d, xb, P, ws, splits, TC_BF16)
        tmp = torch.empty((Cx, Kp), dtype=torch.float32, device=dy.device)
        _lib.K.wgrad_reduce(ws, splits, Cx, 1, 1, Kp, tmp, 0, Kp, 0, 0, 1)
        dw = torch.empty_like(w)
        dw.copy_(tmp[:, :K].view(Cx, g.kh, g.kw, Cy).permute(3, 0, 1, 2))     # [ci][(ky,kx,co)] -> (co, ci, ky, kx)
    return dx, dw


def conv_forward(g: ConvGeom, packs: WeightPacks, w, x, x_layout, out_layout, bias=None, scale=None, relu=False,
                 out_dtype=None, stats_box=None):
    """Y = epilogue(conv(X, W)); the tcgen05 path takes a bf16 activation (cast here if it is not), the fp32 path either"""
    N, Hx, Wx, Cx, xs = _dims(x, x_layout)
    assert Cx == g.Cx, (Cx, g.Cx)
    Hy, Wy = g.out_hw(Hx, Wx)
    tc = _tc_fwd_ok(g, x_layout)
    odt = _out_dtype(x, x_layout, out_dtype)
    x = tc_operand(x, tc)
    wmat, ldw = packs.get(("fwd", tc) + g.key(), w, lambda src, rec: _pack_fwd(g, src, tc, rec))
    y, ys = _empty(N, Hy, Wy, g.Cy, out_layout, x.device, odt)
    d = ConvDesc(B=N, Qh=Hy, Qw=Wy, Cin=g.Cx, Cout=g.Cy, Th=g.kh, Tw=g.kw, in_sy=g.s, in_sx=g.s, tap_sy=1, tap_sx=1,
                 tap_oy=-g.p, tap_ox=-g.p, Hi=Hx, Wi=Wx, up_shift=0, in_sn=xs[0], in_sh=xs[1], in_sw=xs[2], in_sc=xs[3],
                 out_sy=1, out_sx=1, out_oy=0, out_ox=0, Ho=Hy, Wo=Wy, out_sn=ys[0], out_sh=ys[1], out_sw=ys[2],
                 out_sc=ys[3], ldw=ldw, relu=int(relu), scale_rows=_scale_rows(scale, N * Hy * Wy))
    st = None
    if stats_box is not None and FUSE_BN_STATS and tc and out_layout == "cl" and _lib.K.conv_stats_ok(d, tc):
        st = _lib.K.conv_stats_buffer(d, x.device)          # (sum, sum^2) per 32-row slab and channel, filled by the epilogue
        stats_box.append(st)
    _lib.K.conv_gemm(d, x, wmat, bias, scale, y, tc, stats=st)
    return y


# When the fork / join levers below (and those of b200gan.step) apply: "capture" = only while the current stream is being
# captured into a CUDA graph (the replayed iteration is GPU bound and gains 5 ms from them; an EAGER iteration is bound by the
# host's ~1900 launches and the extra stream switches only cost it time: 46 -> 57-65 ms), "always", or "off".
FORKS = os.environ.get("B200_FORKS", "capture")


def forks_enabled(device=None) -> bool:
    if FORKS == "always":
        return True
    return FORKS == "capture" and torch.cuda.is_current_stream_capturing()


SIDE_WGRAD = os.environ.get("B200_SIDE_WGRAD", "1") != "0"
SIDE_WGRAD_MAX_TILES = int(os.environ.get("B200_SIDE_WGRAD_MAX_TILES", "444"))
PARALLEL_PHASES = os.environ.get("B200_PARALLEL_PHASES", "1") != "0"
PARALLEL_PHASES_MAX_TILES = int(os.environ.get("B200_PARALLEL_PHASES_MAX_TILES", "444"))      # per phase; 3 CTAs x 148 SMs
PER_STREAM_FORKS = os.environ.get("B200_PER_STREAM_FORKS", "1") != "0"
_PHASE_STREAMS: Dict[tuple, list] = {}


def _phase_streams(device, n: int):
    """side streams of the CURRENT stream (the three discriminators run on their own streams, TrainStep._side_by_side: each
    gets its own set, so their forked weight-gradient chains do not serialise on one shared stream)"""
    cur = torch.cuda.current_stream(device).cuda_stream if PER_STREAM_FORKS else 0
    pool = _PHASE_STREAMS.setdefault((str(device), cur), [])
    while len(pool) < n:
        pool.append(torch.cuda.Stream(device=device))
    return pool[:n]


# measured: forking the discriminator blocks' shortcut branch costs 0.5 ms per step (30.3 -> 30.85: the pooling pass and the
# small 1x1 GEMM then compete with the residual branch's full-width GEMMs instead of filling a gap) — off
FORK_BRANCHES = os.environ.get("B200_FORK_BRANCHES", "0") != "0"


class forked:
    """`with forked(x) as fk: ...` runs the body on a side stream of the current stream (a no-op off the GPU); `fk.join()`
    makes the forking stream wait for it.  For independent branches of a module (a discriminator block's shortcut): autograd
    keeps every node on its forward stream, so the branch's backward overlaps as well.  `x` (any tensor the branch reads) names
    the device."""

    def __init__(self, ref: torch.Tensor, index: int = 4):
        self.on = FORK_BRANCHES and ref.is_cuda and forks_enabled()
        if self.on:
            self.cur = torch.cuda.current_stream(ref.device)
            self.side = _phase_streams(ref.device, index + 1)[index]

    def __enter__(self):
        if self.on:
            self.side.wait_stream(self.cur)
            self.ctx = torch.cuda.stream(self.side)
            self.ctx.__enter__()
        return self

    def __exit__(self, *exc):
        if self.on:
            self.ctx.__exit__(*exc)
        return False

    def join(self):
        if self.on:
            self.cur.wait_stream(self.side)


_SIDE_STREAMS: Dict[str, list] = {}


def side_streams(device, n: int):
    """streams for coarse fork / join regions (TrainStep._side_by_side); distinct from the per-node phase / wgrad streams"""
    pool = _SIDE_STREAMS.setdefault(str(device), [])
    while len(pool) < n:
        pool.append(torch.cuda.Stream(device=device))
    return pool[:n]


def all_forked_streams(device):
    """the streams autograd nodes may live on besides the caller's (the coarse fork / join regions of TrainStep): a collective
    over gradients produced there waits for these (b200gan.ddp).  The per-node phase / weight-gradient streams are not listed:
    they join their forking stream before the node returns."""
    return list(_SIDE_STREAMS.get(str(device), []))


def conv_dgrad(g: ConvGeom, packs: WeightPacks, w, dy, dy_layout, x_hw, out_layout, scale=None, out_dtype=None, mask=None):
    """dX = conv^T(dY, W)  (also the forward of nn.ConvTranspose2d).  mask: the layer input X when X is a ReLU output whose
    producer expects the masked gradient — dX is zeroed where X <= 0 (in the GEMM epilogue on the tcgen05 path)"""
    N, Hy, Wy, Cy, ds = _dims(dy, dy_layout)
    assert Cy == g.Cy, (Cy, g.Cy)
    Hx, Wx = x_hw
    tc = _tc_dgrad_ok(g, dy_layout)
    odt = _out_dtype(dy, dy_layout, out_dtype)
    dy = tc_operand(dy, tc)
    phase_packs = packs.get(("dgrad", tc) + g.key(), w, lambda src, rec: _pack_dgrad(g, src, tc, rec))
    phases = _dgrad_phases(g)
    dx, xs = _empty(N, Hx, Wx, g.Cx, out_layout, dy.device, odt)
    if any(p is None for p in phase_packs):
        dx.zero_()
    launches = []
    for (py, px, ky0, kx0, Th, Tw), pk in zip(phases, phase_packs):
        if pk is None:
            continue
        Qh, Qw = (Hx - py + g.s - 1) // g.s, (Wx - px + g.s - 1) // g.s
        if Qh <= 0 or Qw <= 0:
            continue
        wmat, ldw = pk
        d = ConvDesc(B=N, Qh=Qh, Qw=Qw, Cin=g.Cy, Cout=g.Cx, Th=Th, Tw=Tw, in_sy=1, in_sx=1, tap_sy=-1, tap_sx=-1,
                     tap_oy=(py + g.p - ky0) // g.s, tap_ox=(px + g.p - kx0) // g.s, Hi=Hy, Wi=Wy, up_shift=0,
                     in_sn=ds[0], in_sh=ds[1], in_sw=ds[2], in_sc=ds[3], out_sy=g.s, out_sx=g.s, out_oy=py, out_ox=px,
                     Ho=Hx, Wo=Wx, out_sn=xs[0], out_sh=xs[1], out_sw=xs[2], out_sc=xs[3], ldw=ldw, relu=0,
                     scale_rows=_scale_rows(scale, N * Qh * Qw))
        launches.append((d, wmat))
    emask = mask if (tc and mask is not None and mask.dtype == dx.dtype) else None
    # the stride^2 output phases are independent launches over the same gradient (disjoint output pixels): small ones — fewer
    # tiles each than the persistent kernel has CTA slots — run side by side on forked streams and join before returning
    side = None
    if PARALLEL_PHASES and tc and dy.is_cuda and len(launches) > 1 and forks_enabled():
        d0 = launches[0][0]
        tiles = -(-(N * d0.Qh * d0.Qw) // 128) * -(-g.Cx // (128 if g.Cx >= 128 else 64))
        if tiles <= PARALLEL_PHASES_MAX_TILES:
            side = _phase_streams(dy.device, len(launches) - 1)
    if side is None:
        for d, wmat in launches:
            _lib.K.conv_gemm(d, dy, wmat, None, scale, dx, tc, emask)
    else:
        cur = torch.cuda.current_stream(dy.device)
        for i, (d, wmat) in enumerate(launches):
            if i == 0:
                _lib.K.conv_gemm(d, dy, wmat, None, scale, dx, tc, emask)
                continue
            st = side[i - 1]
            st.wait_stream(cur)
            with torch.cuda.stream(st):
                _lib.K.conv_gemm(d, dy, wmat, None, scale, dx, tc, emask)
        for st in side:
            cur.wait_stream(st)
    if mask is not None and not (tc and mask.dtype == dx.dtype):
        dx = _lib.K.relu_bwd(dx, mask.to(dx.dtype) if mask.dtype != dx.dtype else mask)
    return dx


def _wgrad_splits(g: ConvGeom, Q: int, tc: bool, rows=None, cols=None) -> int:
    if tc:
        tiles = -(-g.Cy // 128) * g.kh * g.kw * -(-g.Cx // (128 if g.Cx >= 128 else 64))
    elif rows is not None:          # flattened (tap, channel) column tiles of the skinny-operand kernel
        tiles = -(-rows // 64) * -(-cols // 64)
    else:
        tiles = -(-g.Cy // 64) * g.kh * g.kw * -(-g.Cx // 64)
    # tcgen05 kernel: one wave of two co-resident CTAs per SM; more splits only add partial-result traffic (every split
    # writes and the reduction re-reads a full fp32 copy of the gradient)
    target = _WGRAD_TARGET if tc else 592
    splits = max(1, min(-(-target // tiles), max(1, Q // 256), 256 if rows is not None else 64))
    return splits


def conv_wgrad(g: ConvGeom, x, x_layout, dy, dy_layout, dw: torch.Tensor, accumulate=False):
    """dW[co, cx_offset+c, ky, kx] (+)= sum X (*) dY; dw is the full (Cy, cx_total, kh, kw) gradient buffer."""
    N, Hx, Wx, Cx, xs = _dims(x, x_layout)
    _, Hy, Wy, Cy, ds = _dims(dy, dy_layout)
    assert Cx == g.Cx and Cy == g.Cy
    tc = _tc_wgrad_ok(g, x_layout, dy_layout)
    x, dy = tc_operand(x, tc), tc_operand(dy, tc)
    kk = g.kh * g.kw
    if not tc and g.Cy <= 8 and g.Cx > 8 and g.s == 1:
        # skinny OUTPUT side (the 64->3 image convolutions): swap the roles so the 3-channel tensor is the gathered,
        # tap-flattened operand:  dW[co,c,ky,kx] = sum_{q'} X[q'][c] * dY[q' - (ky,kx) + p][co]
        Q = N * Hx * Wx
        K = kk * g.Cy
        splits = _wgrad_splits(g, Q, False, rows=g.Cx, cols=K)
        ws = torch.empty((splits * g.Cx * K,), dtype=torch.float32, device=x.device)
        d = ConvDesc(B=N, Qh=Hx, Qw=Wx, Cin=g.Cy, Cout=g.Cx, Th=g.kh, Tw=g.kw, in_sy=1, in_sx=1, tap_sy=-1, tap_sx=-1,
                     tap_oy=g.p, tap_ox=g.p, Hi=Hy, Wi=Wy, up_shift=0, in_sn=ds[0], in_sh=ds[1], in_sw=ds[2], in_sc=ds[3],
                     out_sy=1, out_sx=1, out_oy=0, out_ox=0, Ho=Hx, Wo=Wx, out_sn=xs[0], out_sh=xs[1], out_sw=xs[2],
                     out_sc=xs[3], ldw=K, relu=0)
        _lib.K.wgrad_gemm(d, x, dy, ws, splits, False)
        # ws rows = input channel c, columns = (ky, kx, co)  ->  dw[co, cx_offset + c, ky, kx]
        _lib.K.wgrad_reduce(ws, splits, g.Cx, g.kh, g.kw, g.Cy, dw, g.cx_offset * kk, kk, g.kw, 1, g.cx_total * kk,
                            accumulate=accumulate)
        return dw
    Q = N * Hy * Wy
    K = kk * g.Cx
    splits = _wgrad_splits(g, Q, tc, rows=(g.Cy if (not tc and g.Cx <= 8) else None), cols=K)
    d = ConvDesc(B=N, Qh=Hy, Qw=Wy, Cin=g.Cx, Cout=g.Cy, Th=g.kh, Tw=g.kw, in_sy=g.s, in_sx=g.s, tap_sy=1, tap_sx=1,
                 tap_oy=-g.p, tap_ox=-g.p, Hi=Hx, Wi=Wx, up_shift=0, in_sn=xs[0], in_sh=xs[1], in_sw=xs[2], in_sc=xs[3],
                 out_sy=1, out_sx=1, out_oy=0, out_ox=0, Ho=Hy, Wo=Wy, out_sn=ds[0], out_sh=ds[1], out_sw=ds[2],
                 out_sc=ds[3], ldw=K, relu=0)
    if tc == TC_TF32:
        # fp32 operands on the bf16 tensor-core kernel: x = xh + xl, dy = dh + dl (two-term bf16 splits, 16 mantissa bits);
        # dY^T X ~ dh^T xh + dh^T xl + dl^T xh, three launches whose partials are summed by the same fixed-order reduction
        n_el = g.Cy * K
        ws = torch.empty((3 * splits * n_el,), dtype=torch.float32, device=x.device)
        xh, xl = _lib.K.split_bf16(x.contiguous())
        dh, dl = _lib.K.split_bf16(dy.contiguous())
        for i, (a, b) in enumerate(((dh, xh), (dh, xl), (dl, xh))):
            _lib.K.wgrad_gemm(d, a, b, ws[i * splits * n_el:(i + 1) * splits * n_el], splits, TC_BF16)
        splits *= 3
    else:
        ws = torch.empty((splits * g.Cy * K,), dtype=torch.float32, device=x.device)
        _lib.K.wgrad_gemm(d, dy, x, ws, splits, tc)
    _lib.K.wgrad_reduce(ws, splits, g.Cy, g.kh, g.kw, g.Cx, dw, g.cx_offset * kk, g.cx_total * kk, g.kw, 1, kk,
                        accumulate=accumulate)
    return dw


GROUPED_SN_WGRAD = True      # one weight-gradient GEMM for all batched calls of a spectral-normalised layer


def _sn_group_splits(g: ConvGeom, Q: int, groups: int) -> int:
    """pixel splits PER CALL for the grouped spectral-norm weight gradient (0 = not applicable): the calls' row ranges must
    be whole numbers of 64-pixel k-blocks of equal size"""
    if not GROUPED_SN_WGRAD or _PRECISION != "bf16" or groups < 2 or groups > 8 or Q % groups:
        return 0
    Qg = Q // groups
    kk = g.kh * g.kw
    if kk > 64 or g.Cy >= 65536 or g.Cy * g.Cx * kk >= 2 ** 31 or g.cx_offset != 0 or g.cx_total != g.Cx:
        return 0
    spg = max(1, _wgrad_splits(g, Q, TC_BF16) // groups)
    while spg > 1 and Qg % (64 * spg):
        spg -= 1
    return spg if Qg % (64 * spg) == 0 else 0


def conv_wgrad_sn_grouped(g: ConvGeom, x, dy, sn: "SNCall", w, spg: int, kind: int = TC_BF16) -> torch.Tensor:
    """dW through W / sigma_g for sn.groups batched calls (x, dy: operands of `kind`, channel-last, rows of call g contiguous): one
    tcgen05 weight-gradient launch with call-aligned pixel splits, then b200_sn_wgrad_finish"""
    N, Hx, Wx, Cx, xs = _dims(x, "cl")
    _, Hy, Wy, Cy, ds = _dims(dy, "cl")
    kk = g.kh * g.kw
    K = kk * g.Cx
    splits = sn.groups * spg
    ws = torch.empty((splits * g.Cy * K,), dtype=torch.float32, device=x.device)
    d = ConvDesc(B=N, Qh=Hy, Qw=Wy, Cin=g.Cx, Cout=g.Cy, Th=g.kh, Tw=g.kw, in_sy=g.s, in_sx=g.s, tap_sy=1, tap_sx=1,
                 tap_oy=-g.p, tap_ox=-g.p, Hi=Hx, Wi=Wx, up_shift=0, in_sn=xs[0], in_sh=xs[1], in_sw=xs[2], in_sc=xs[3],
                 out_sy=1, out_sx=1, out_oy=0, out_ox=0, Ho=Hy, Wo=Wy, out_sn=ds[0], out_sh=ds[1], out_sw=ds[2],
                 out_sc=ds[3], ldw=K, relu=0)
    _lib.K.wgrad_gemm(d, dy, x, ws, splits, kind)
    dw = torch.empty_like(w)
    return _lib.K.sn_wgrad_finish(ws, sn.groups, spg, g.Cy, kk, g.Cx, w, sn.u_hist, sn.v_hist, sn.inv, dw,
                                  pooled_taps=(g.kh - 1, g.kw - 1) if g.fold else None)


_BIAS_GRAD_MEMO: List[tuple] = []      # the last few (weak reference to a gradient tensor, its column sums)


def bias_grad(dy: torch.Tensor, layout: str) -> torch.Tensor:
    """column sums of an output gradient.  Two convolutions that receive the SAME gradient tensor (a discriminator block's
    second convolution and its shortcut: pool(a + b) hands one tensor to both branches) share one reduction — the memo holds a
    weak references to the last three tensors, so a hit is only possible while that very tensor object is alive."""
    if layout == "cl":
        for ref, val in _BIAS_GRAD_MEMO:
            if ref() is dy and val.shape[0] == dy.shape[-1]:
                return val.clone()
        val = _lib.K.colsum(dy.reshape(-1, dy.shape[-1]))
        _BIAS_GRAD_MEMO.insert(0, (weakref.ref(dy), val))
        del _BIAS_GRAD_MEMO[3:]
        return val
    N, C, H, W = dy.shape
    per = _lib.K.rowsum(dy.contiguous().view(N * C, H * W))          # per-(n, c) sums, then over n
    return _lib.K.colsum(per.view(N, C))


# ----------------------------------------------------------------------------------------------------------
# convolution / transposed convolution / linear  (+ optional spectral norm, bias, fused ReLU)
# ----------------------------------------------------------------------------------------------------------
class SNCall:
    """State of one (possibly batched) spectral-norm evaluation of a layer: `groups` sequential power iterations
    (one per batched call, torch.nn.utils.spectral_norm's pre-forward hook run `groups` times) -> inv (groups,) = 1/sigma
    per call and the u, v vectors each call's sigma was computed with (u_hist (groups,h), v_hist (groups,w))."""

    __slots__ = ("groups", "inv", "u_hist", "v_hist")

    def __init__(self, groups, inv, u_hist, v_hist):
        self.groups, self.inv, self.u_hist, self.v_hist = groups, inv, u_hist, v_hist


def sn_iterate(w, u, v, groups: int, training: bool) -> SNCall:
    """`groups` power iterations of one layer, in place on its u / v buffers (discriminator.py:15-22 hook semantics)"""
    h, wd = w.shape[0], w[0].numel()
    inv = torch.empty((groups,), dtype=torch.float32, device=w.device)
    u_hist = torch.empty((groups, h), dtype=torch.float32, device=w.device)
    v_hist = torch.empty((groups, wd), dtype=torch.float32, device=w.device)
    for gi in range(groups):
        _lib.K.sn_power_iter(w, h, wd, u, v, training, SN_EPS, inv_out=inv[gi:gi + 1])
        _lib.K.copy_into(u_hist, gi, u.view(1, h))
        _lib.K.copy_into(v_hist, gi, v.view(1, wd))
    return SNCall(groups, inv, u_hist, v_hist)


class SNPlan:
    """Whole-network power iteration (b200_sn_power_iter_multi): the device descriptor table, scratch and staging buffers
    for the spectral-normalised layers of one network and one number of batched calls."""

    def __init__(self, mods, groups: int):
        self.groups = groups
        self.mods = list(mods)
        self.stamp = self._stamp(self.mods)
        dev = self.mods[0].weight_orig.device
        layers, so, wo = [], 0, 0
        for m in self.mods:
            W = m.weight_orig
            h, wd = W.shape[0], W[0].numel()
            layers.append((W, m.weight_u, m.weight_v, h, wd, so, wo))
            so += groups * (1 + h + wd)
            wo += _rup(8 * wd + h, 4)          # 16-byte aligned scratch per layer (vector stores)
        self.layers = layers
        self.stage = torch.empty((so,), dtype=torch.float32, device=dev)
        self.ws = torch.empty((wo,), dtype=torch.float32, device=dev)
        self.table = _lib.K.sn_table(layers, self.stage, self.ws, groups)

    @staticmethod
    def _stamp(mods):
        return tuple((m.weight_orig.data_ptr(), m.weight_u.data_ptr(), m.weight_v.data_ptr()) for m in mods)

    def valid_for(self, mods) -> bool:
        return self._stamp(mods) == self.stamp

    def run(self, training: bool) -> List[SNCall]:
        """`groups` iterations of every layer (4 launches per iteration), then one copy of the staged results into a
        buffer this call owns (so a later call of the network cannot overwrite what this call's backward needs)."""
        g = self.groups
        _lib.K.sn_power_iter_multi(self.table, self.layers, self.stage, self.ws, g, training, SN_EPS)
        own = torch.empty_like(self.stage)
        _lib.K.copy_into(own.view(1, -1), 0, self.stage.view(1, -1))
        calls = []
        for (_, _, _, h, wd, so, _) in self.layers:
            calls.append(SNCall(g, own[so:so + g], own[so + g:so + g + g * h].view(g, h),
                                own[so + g + g * h:so + g + g * h + g * wd].view(g, wd)))
        return calls


class _ConvFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, bias, sn: Optional[SNCall], g: ConvGeom, packs: WeightPacks, transposed: bool, x_layout: str,
                out_layout: str, relu: bool, out_hw, out_dtype, mask_input_grad: bool = False, grad_premasked: bool = False,
                stats_box=None):
        # mask_input_grad: x is the ReLU output of a producer that set grad_premasked — this layer's data gradient is
        # returned already multiplied by (x > 0); grad_premasked: the incoming gradient of a fused-ReLU layer is already
        # masked by its (single) consumer, so the separate relu_bwd pass is skipped
        ctx.mask_input_grad, ctx.grad_premasked = bool(mask_input_grad), bool(grad_premasked)
        scale = sn.inv if sn is not None else None
        ctx.x_dtype = x.dtype
        ctx.sn = sn
        if not transposed:
            fwd_tc, wgrad_tc = _tc_fwd_ok(g, x_layout), _tc_wgrad_ok(g, x_layout, out_layout)
        else:
            fwd_tc, wgrad_tc = _tc_dgrad_ok(g, x_layout), _tc_wgrad_ok(g, out_layout, x_layout)
        x_op = tc_operand(x, fwd_tc)            # cast once; the bf16 copy is also what a tcgen05 weight gradient consumes
        ctx.packed = (not transposed) and _packed_ok(g) and out_layout == "cl"
        if ctx.packed:
            # few-channel input: explicit bf16 im2col matrix, kept for the weight gradient
            N, Hx, Wx, _, _ = _dims(x, x_layout)
            Hy, Wy = g.out_hw(Hx, Wx)
            x_op = im2col_pack(g, x, x_layout)
            ctx.y_dims = (N, Hy, Wy)
            y = conv_forward_packed(g, packs, w, x_op, N, Hy, Wy, out_layout, bias, scale, relu,
                                    _out_dtype(x, x_layout, out_dtype), stats_box)
        elif not transposed:
            y = conv_forward(g, packs, w, x_op, x_layout, out_layout, bias, scale, relu, _out_dtype(x, x_layout, out_dtype),
                             stats_box)
        else:
            assert bias is None and not relu
            y = conv_dgrad(g, packs, w, x_op, x_layout, out_hw, out_layout, scale, _out_dtype(x, x_layout, out_dtype))
        ctx.g, ctx.packs, ctx.transposed = g, packs, transposed
        ctx.x_layout, ctx.out_layout, ctx.relu = x_layout, out_layout, relu
        ctx.has_bias = bias is not None
        ctx.x_dims = _dims(x, x_layout)[:4]
        ctx.save_for_backward(x_op if ((fwd_tc and wgrad_tc) or ctx.packed) else x, w, y if relu else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, y = ctx.saved_tensors
        g, packs, sn = ctx.g, ctx.packs, ctx.sn
        dy = dy.contiguous()
        if ctx.relu and not ctx.grad_premasked:
            dy = _lib.K.relu_bwd(dy, y)
        scale = sn.inv if sn is not None else None
        dx = dw = db = None
        need_dx, need_dw = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        if not ctx.transposed:
            dgrad_tc, wgrad_tc = _tc_dgrad_ok(g, ctx.out_layout), _tc_wgrad_ok(g, ctx.x_layout, ctx.out_layout)
        else:
            dgrad_tc, wgrad_tc = _tc_fwd_ok(g, ctx.out_layout), _tc_wgrad_ok(g, ctx.out_layout, ctx.x_layout)
        packed = ctx.packed
        if (not ctx.transposed) and sn is None and _packed_out_ok(g, ctx.x_layout) and (need_dx or need_dw):
            dx, dw = conv_backward_packed_out(g, packs, w, x, ctx.x_dims, dy, ctx.out_layout, need_dx, need_dw, scale,
                                              ctx.x_dtype)
            if ctx.has_bias and ctx.needs_input_grad[2]:
                db = bias_grad(dy, ctx.out_layout)
            return dx, dw, db, None, None, None, None, None, None, None, None, None, None, None, None
        # one cast for both GEMMs (the tf32 kernels take the fp32 gradient as it is)
        dyb = tc_operand(dy, TC_BF16 if packed else max(dgrad_tc, wgrad_tc)) \
            if ((need_dx and dgrad_tc) or (need_dw and (wgrad_tc or packed))) else None
        # small layers (fewer output tiles than the GPU has CTA slots): the weight-gradient chain (GEMM, split reduction,
        # spectral-norm finish) runs on a forked stream next to the data-gradient launches and joins before this node returns
        side = None
        if SIDE_WGRAD and need_dx and need_dw and dy.is_cuda and dgrad_tc and wgrad_tc and forks_enabled():
            rows = dy.numel() // dy.shape[-1] if ctx.out_layout == "cl" else 0
            if 0 < -(-rows // 128) * -(-g.Cy // (128 if g.Cy >= 128 else 64)) <= SIDE_WGRAD_MAX_TILES:
                cur = torch.cuda.current_stream(dy.device)
                side = _phase_streams(dy.device, 4)[3]
                side.wait_stream(cur)
        if need_dx:
            dy_op = dyb if dgrad_tc else dy
            if not ctx.transposed:
                _, Hx, Wx, _ = ctx.x_dims
                dx = conv_dgrad(g, packs, w, dy_op, ctx.out_layout, (Hx, Wx), ctx.x_layout, scale, ctx.x_dtype,
                                mask=x if (ctx.mask_input_grad and not ctx.packed) else None)
            else:
                dx = conv_forward(g, packs, w, dy_op, ctx.out_layout, ctx.x_layout, None, scale, False, ctx.x_dtype)
        if need_dw:
          with (torch.cuda.stream(side) if side is not None else contextlib.nullcontext()):
              gw = torch.empty_like(w)
              dy_op = dyb if (wgrad_tc or packed) else dy

              def wgrad(xx, xl, dd, dl):
                  """the plain weight gradient in the parameter's shape (a pooled convolution: through its folded image)"""
                  if not g.fold:
                      return conv_wgrad(g, xx, xl, dd, dl, gw)
                  g4 = torch.empty((g.Cy, g.cx_total, g.kh, g.kw), dtype=torch.float32, device=w.device)
                  return unfold_pool_grad(conv_wgrad(g, xx, xl, dd, dl, g4))

              assert not (g.fold and (packed or ctx.transposed))
              if packed:
                  N, Hy, Wy = ctx.y_dims
                  if sn is None:
                      dw = conv_wgrad_packed(g, x, N, Hy, Wy, dy_op, ctx.out_layout, gw)
                  else:
                      n = N // sn.groups
                      rows = n * Hy * Wy
                      h, wd = w.shape[0], w[0].numel()
                      dw = torch.empty_like(w)
                      for gi in range(sn.groups):
                          conv_wgrad_packed(g, x[gi * rows:(gi + 1) * rows], n, Hy, Wy, dy_op[gi * n:(gi + 1) * n],
                                            ctx.out_layout, gw)
                          _lib.K.sn_grad(gw, w, sn.u_hist[gi], sn.v_hist[gi], sn.inv[gi:gi + 1], h, wd, dW=dw, accumulate=gi > 0)
              elif sn is None:
                  if not ctx.transposed:
                      dw = wgrad(x, ctx.x_layout, dy_op, ctx.out_layout)
                  else:
                      dw = wgrad(dy_op, ctx.out_layout, x, ctx.x_layout)
              elif wgrad_tc and _sn_group_splits(g, dy_op.shape[0] * dy_op.shape[1] * dy_op.shape[2], sn.groups) > 0:
                  assert not ctx.transposed
                  spg = _sn_group_splits(g, dy_op.shape[0] * dy_op.shape[1] * dy_op.shape[2], sn.groups)
                  dw = conv_wgrad_sn_grouped(g, x, dy_op, sn, w, spg, wgrad_tc)
              else:
                  # batched calls have their own sigma, u, v: gradient through W / sigma_g per group of rows
                  assert not ctx.transposed
                  n = x.shape[0] // sn.groups
                  h, wd = w.shape[0], w[0].numel()
                  dw = torch.empty_like(w)
                  for gi in range(sn.groups):
                      gg = wgrad(x[gi * n:(gi + 1) * n], ctx.x_layout, dy_op[gi * n:(gi + 1) * n], ctx.out_layout)
                      _lib.K.sn_grad(gg, w, sn.u_hist[gi], sn.v_hist[gi], sn.inv[gi:gi + 1], h, wd, dW=dw, accumulate=gi > 0)
        if side is not None:
            cur.wait_stream(side)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = bias_grad(dy, ctx.out_layout)
        return dx, dw, db, None, None, None, None, None, None, None, None, None, None, None, None


def conv2d(x, w, bias, g: ConvGeom, packs: WeightPacks, x_layout="cl", out_layout="cl", relu=False,
           sn: Optional[SNCall] = None, out_dtype=None, mask_input_grad=False, grad_premasked=False, stats=False):
    """stats: the output feeds a batch norm — ask the kernel's epilogue for the per-slab (sum, sum of squares) pairs; when the
    launch can provide them they ride on the returned tensor (`_b200_colstats`) and the normalisation skips its statistics pass"""
    box = [] if stats else None
    y = _ConvFn.apply(x, w, bias, sn, g, packs, False, x_layout, out_layout, relu, None, out_dtype, mask_input_grad,
                      grad_premasked, box)
    if box:
        y._b200_colstats = box[0]
    return y


def conv_transpose2d(x, w, g: ConvGeom, packs: WeightPacks, out_hw, x_layout="cl", out_layout="cl"):
    """x plays dY of the conv-orientation geometry g (g.Cy = x channels, g.Cx = output channels)."""
    return _ConvFn.apply(x, w, None, None, g, packs, True, x_layout, out_layout, False, out_hw, None)


def linear(x2d, w, bias, packs: WeightPacks, relu=False, sn: Optional[SNCall] = None, geom: Optional[ConvGeom] = None,
           out_dtype=None):
    B, Cin = x2d.shape
    g = geom if geom is not None else ConvGeom(Cin, w.shape[0], 1, 1, 1, 0)
    y = _ConvFn.apply(x2d.view(B, 1, 1, Cin), w.view(w.shape[0], w.shape[1], 1, 1), bias, sn, g, packs, False, "cl",
                      "cl", relu, None, out_dtype)
    return y.view(B, g.Cy)


# ----------------------------------------------------------------------------------------------------------
# normalisation
# ----------------------------------------------------------------------------------------------------------
# Synchronised batch statistics (the mechanism of the reference's vendored
# models/spade/networks/sync_batchnorm/batchnorm.py:74-145): under data parallelism every normalisation layer all-reduces its
# per-channel (sum x, sum x^2, count) in the forward pass and its (sum dxhat, sum dxhat * xhat) in the backward pass, so the
# sharded step computes exactly the statistics — and, with the count-weighted losses of TrainStep, exactly the gradients — of
# the single-process global batch.  Off by default (DDP semantics: per-shard statistics).
_SYNC_BN = None          # process group (or True for the default group) while enabled


def set_sync_bn(group):
    """group: a torch.distributed process group, True for the default group, None to disable"""
    global _SYNC_BN
    _SYNC_BN = group


def _sync_group():
    return None if _SYNC_BN is True else _SYNC_BN


def _sync_moments(mean, var, n_local, running_mean, running_var):
    """global (mean, biased var) per group from the shards' local ones; running statistics updated with the GLOBAL values in
    call order (momentum 0.1, unbiased variance).  Returns mean, var (fp32) and n_local / n_global per group (G, 1) fp32."""
    import torch.distributed as dist
    G, C = mean.shape
    buf = torch.empty((G, 2 * C + 1), dtype=torch.float64, device=mean.device)
    m64 = mean.double()
    buf[:, :C] = m64 * n_local
    buf[:, C:2 * C] = (var.double() + m64 * m64) * n_local
    buf[:, 2 * C] = float(n_local)
    dist.all_reduce(buf, group=_sync_group())
    n = buf[:, 2 * C:2 * C + 1]
    gm = buf[:, :C] / n
    gv = (buf[:, C:2 * C] / n - gm * gm).clamp_(min=0.0)
    if running_mean is not None:
        unb = gv * (n / (n - 1.0).clamp(min=1.0))
        for g in range(G):
            running_mean.mul_(1.0 - BN_MOMENTUM).add_(gm[g].float(), alpha=BN_MOMENTUM)
            running_var.mul_(1.0 - BN_MOMENTUM).add_(unb[g].float(), alpha=BN_MOMENTUM)
    return gm.float().contiguous(), gv.float().contiguous(), (float(n_local) / n).float()


class _NormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, idx, running_mean, running_var, mode, training, relu, residual, rows_per_seg,
                num_classes, groups, colstats=None):
        shape = x.shape
        C = shape[-1]
        x2 = x.reshape(-1, C)
        ctx.sync_ratio = None
        if (training and _SYNC_BN is None and colstats is not None and (x2.shape[0] // groups) % 32 == 0
                and colstats.shape[0] * 32 >= x2.shape[0] and colstats.shape[1] >= C):
            # statistics already accumulated by the producing convolution's epilogue: finish them without re-reading x
            mean, var = _lib.K.bn_stats_slabs(colstats, x2.shape[0], C, running_mean, running_var, BN_MOMENTUM, groups)
        elif training and _SYNC_BN is not None:
            mean, var = _lib.K.bn_stats(x2, None, None, BN_MOMENTUM, groups)
            mean, var, ctx.sync_ratio = _sync_moments(mean, var, x2.shape[0] // groups, running_mean, running_var)
        elif training:
            mean, var = _lib.K.bn_stats(x2, running_mean, running_var, BN_MOMENTUM, groups)
        else:
            mean, var, groups = running_mean, running_var, 1
        ctx.groups = groups
        g2 = gamma.reshape(-1, 2 * C) if mode == MODE_SPADE else gamma
        r2 = residual.reshape(-1, C) if residual is not None else None
        y = _lib.K.norm_fwd(x2, mean, var, BN_EPS, mode, g2, beta, idx, rows_per_seg, r2, relu, groups)
        # conditional batch norm + ReLU: the backward recomputes the ReLU mask from x (gamma * xhat + beta > 0) instead of
        # reading the saved output (relu flag 2; one activation read less per backward pass, y is not kept alive)
        recompute = bool(relu) and mode == MODE_CBN and residual is None and _lib.vec_layout_ok(x2)
        relu_flag = 2 if recompute else int(bool(relu))
        ctx.mode, ctx.relu, ctx.rows_per_seg, ctx.num_classes, ctx.training = mode, relu_flag, rows_per_seg, num_classes, training
        ctx.has_residual = residual is not None
        ctx.save_for_backward(x2, y if relu_flag == 1 else None, mean, var, g2, idx)
        ctx.shape = shape
        return y.view(shape)

    @staticmethod
    def backward(ctx, dy):
        if not ctx.training:
            raise NotImplementedError("b200gan: backward through eval-mode batch norm is not on the reference path")
        x2, y, mean, var, g2, idx = ctx.saved_tensors
        C = x2.shape[1]
        dy2 = dy.contiguous().reshape(-1, C)
        sync = None
        if ctx.sync_ratio is not None:
            ratio, grp, G = ctx.sync_ratio, _sync_group(), ctx.groups

            def sync(s_):
                # (sum dxhat, sum dxhat * xhat) summed over the ranks; the apply kernel divides by the LOCAL row count, so the
                # global mean is sum / n_global = (sum * n_local / n_global) / n_local
                import torch.distributed as dist
                dist.all_reduce(s_, group=grp)
                return (s_.view(G, -1) * ratio).reshape(-1).contiguous()
        dx, dgamma, dbeta, dtable, dgb = _lib.K.norm_bwd(dy2, x2, y, mean, var, BN_EPS, ctx.mode, g2, idx,
                                                         ctx.rows_per_seg, ctx.relu, ctx.num_classes, ctx.groups, sync=sync)
        dres = None
        if ctx.has_residual:
            dres = _lib.K.relu_bwd(dy2, y).view(ctx.shape) if ctx.relu else dy
        if ctx.mode == MODE_AFFINE:
            gg, gb = dgamma, dbeta
        elif ctx.mode == MODE_CBN:
            gg, gb = dtable, None
        elif ctx.mode == MODE_SPADE:
            gg, gb = dgb.view(ctx.shape[:-1] + (2 * C,)), None
        else:
            gg, gb = None, None
        return dx.view(ctx.shape), gg, gb, None, None, None, None, None, None, dres, None, None, None, None


def batch_norm(x, weight, bias, running_mean, running_var, training, relu=False, residual=None, groups=1):
    """groups > 1: `groups` calls of the layer batched along dim 0, each with its own batch statistics"""
    mode = MODE_AFFINE if weight is not None else MODE_PLAIN
    return _NormFn.apply(x, weight, bias, None, running_mean, running_var, mode, training, relu, residual, 1, 0, groups,
                         getattr(x, "_b200_colstats", None))


def cond_batch_norm(x, table, idx_i32, running_mean, running_var, training, relu=False, groups=1):
    """x (O,H,W,C); table (num_classes, 2C) = [gamma | beta]; idx_i32 (O,)"""
    rows_per_seg = x.shape[1] * x.shape[2]
    return _NormFn.apply(x, table, None, idx_i32, running_mean, running_var, MODE_CBN, training, relu, None,
                         rows_per_seg, table.shape[0], groups, getattr(x, "_b200_colstats", None))


def spade_norm(x, gb, running_mean, running_var, training, relu=False, groups=1):
    """x (N,H,W,C); gb (N,H,W,2C) = fused [gamma | beta] conv output"""
    return _NormFn.apply(x, gb, None, None, running_mean, running_var, MODE_SPADE, training, relu, None, 1, 0, groups,
                         getattr(x, "_b200_colstats", None))


# ----------------------------------------------------------------------------------------------------------
# pointwise / pooling / layout
# ----------------------------------------------------------------------------------------------------------
class _ReluFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        y = _lib.K.relu_fwd(x.contiguous())
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        return _lib.K.relu_bwd(dy.contiguous(), y)


class _AddFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        return _lib.K.add(a.contiguous(), b.contiguous())

    @staticmethod
    def backward(ctx, dy):
        return dy, dy


class _PoolFn(torch.autograd.Function):
    """y = scale * sum over f x f blocks (cl tensors; an NCHW tensor is passed as (N*C, H, W, 1))"""

    @staticmethod
    def forward(ctx, x, f, scale):
        N, H, W, C = x.shape
        ctx.f, ctx.scale = f, scale
        return _lib.K.pool_fwd(x.contiguous(), N, H, W, C, f, scale)

    @staticmethod
    def backward(ctx, dy):
        N, H, W, C = dy.shape
        return _lib.K.unpool_fwd(dy.contiguous(), N, H, W, C, ctx.f, ctx.scale), None, None


class _PoolAddFn(torch.autograd.Function):
    """y = [relu](scale * (f x f block sums of a + b)); both inputs receive the same gradient tensor"""

    @staticmethod
    def forward(ctx, a, b, f, scale, relu):
        N, H, W, C = a.shape
        ctx.f, ctx.scale, ctx.relu = f, scale, relu
        y = _lib.K.pool_add_fwd(a.contiguous(), b.contiguous(), N, H, W, C, f, scale, relu)
        if relu:
            ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, dy):
        N, H, W, C = dy.shape
        if ctx.relu:
            (y,) = ctx.saved_tensors
            g = _lib.K.unpool_masked_fwd(dy.contiguous(), y, N, H, W, C, ctx.f, ctx.scale)
        else:
            g = _lib.K.unpool_fwd(dy.contiguous(), N, H, W, C, ctx.f, ctx.scale)
        return g, g, None, None, None


def avg_pool2_sum(a, b, relu=False):
    """[relu](avg_pool2(a + b)) in one pass"""
    return _PoolAddFn.apply(a, b, 2, 0.25, relu)


class _AddReluFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        y = _lib.K.add_relu(a.contiguous(), b.contiguous())
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        g = _lib.K.relu_bwd(dy.contiguous(), y)
        return g, g


def add_relu(a, b):
    return _AddReluFn.apply(a, b)


class _UnpoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, f, scale):
        N, H, W, C = x.shape
        ctx.f, ctx.scale = f, scale
        return _lib.K.unpool_fwd(x.contiguous(), N, H, W, C, f, scale)

    @staticmethod
    def backward(ctx, dy):
        N, H, W, C = dy.shape
        return _lib.K.pool_fwd(dy.contiguous(), N, H, W, C, ctx.f, ctx.scale), None, None


def relu(x):
    return _ReluFn.apply(x)


def add(a, b):
    return _AddFn.apply(a, b)


def avg_pool2(x):
    return _PoolFn.apply(x, 2, 0.25)


def pool(x, f, scale):
    return _PoolFn.apply(x, f, scale)


def upsample_nearest(x, f):
    return _UnpoolFn.apply(x, f, 1.0)


def pool_nchw(x, f, scale):
    N, C, H, W = x.shape
    return _PoolFn.apply(x.reshape(N * C, H, W, 1), f, scale).view(N, C, H // f, W // f)


def upsample_nearest_nchw(x, f):
    N, C, H, W = x.shape
    return _UnpoolFn.apply(x.reshape(N * C, H, W, 1), f, 1.0).view(N, C, H * f, W * f)


class _TransposeFn(torch.autograd.Function):
    """(B, R, C) -> (B, C, R)"""

    @staticmethod
    def forward(ctx, x):
        B, R, C = x.shape
        return _lib.K.transpose(x.contiguous(), B, R, C)

    @staticmethod
    def backward(ctx, dy):
        B, C, R = dy.shape
        return _lib.K.transpose(dy.contiguous(), B, C, R)


def nchw_to_cl(x):
    N, C, H, W = x.shape
    return _TransposeFn.apply(x.reshape(N, C, H * W)).view(N, H, W, C)


def cl_to_nchw(x):
    N, H, W, C = x.shape
    return _TransposeFn.apply(x.reshape(N, H * W, C)).view(N, C, H, W)


class _ConcatFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, a_div, b_div, rows):
        Ca, Cb = a.shape[-1], b.shape[-1]
        ctx.meta = (Ca, a_div, Cb, b_div, rows, a.shape, b.shape)
        return _lib.K.concat_fwd(a.contiguous(), Ca, a_div, b.contiguous(), Cb, b_div, rows)

    @staticmethod
    def backward(ctx, dout):
        Ca, a_div, Cb, b_div, rows, sa, sb = ctx.meta
        da, db = _lib.K.concat_bwd(dout.contiguous(), Ca, a_div, Cb, b_div, rows, ctx.needs_input_grad[0],
                                   ctx.needs_input_grad[1])
        return (da.view(sa) if da is not None else None), (db.view(sb) if db is not None else None), None, None, None


def concat_channels(a, b, a_div=1, b_div=1, rows=None):
    """out[row] = [a[row // a_div] | b[row // b_div]]; a, b are (rows/div, C) views"""
    if rows is None:
        rows = a.shape[0] * a_div
    return _ConcatFn.apply(a, b, a_div, b_div, rows)


class _EmbeddingFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, table, idx_i32):
        ctx.save_for_backward(idx_i32)
        ctx.n = table.shape[0]
        return _lib.K.gather_rows(table, idx_i32)

    @staticmethod
    def backward(ctx, dout):
        (idx,) = ctx.saved_tensors
        return _lib.K.scatter_rows(dout.contiguous(), idx, ctx.n), None


def embedding(table, idx_i32):
    return _EmbeddingFn.apply(table, idx_i32)


class _MaskOuterFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, v, mask):
        O, H, W = mask.shape[0], mask.shape[-2], mask.shape[-1]
        ctx.save_for_backward(mask)
        ctx.dims = (O, H, W, v.shape[1])
        return _lib.K.mask_outer_fwd(v.contiguous(), mask, O, H, W, v.shape[1], act_dtype())

    @staticmethod
    def backward(ctx, dout):
        (mask,) = ctx.saved_tensors
        O, H, W, C = ctx.dims
        return _lib.K.mask_outer_bwd(dout.contiguous(), mask, O, H, W, C), None


def mask_outer(v, mask):
    """(O,C) x (O,1,H,W) -> (O,H+2,W+2,C): the padding-1 1x1 convolution of embedding (x) mask (rank-1 form)"""
    return _MaskOuterFn.apply(v, mask.contiguous())


class _ReparamFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mu, logvar, eps):
        ctx.save_for_backward(logvar, eps)
        return _lib.K.reparam_fwd(mu.contiguous(), logvar.contiguous(), eps)

    @staticmethod
    def backward(ctx, dz):
        logvar, eps = ctx.saved_tensors
        dmu, dlv = _lib.K.reparam_bwd(dz.contiguous(), logvar, eps)
        return dmu, dlv, None


def reparameterize(mu, logvar, eps):
    return _ReparamFn.apply(mu, logvar, eps)


# ----------------------------------------------------------------------------------------------------------
# per-batch index plans (obj_to_img lives on the CPU by contract, train64.py:150)
# ----------------------------------------------------------------------------------------------------------
class BatchPlan:
    """Device-side index tensors derived from obj_to_img: crop grouping and the time-major ConvLSTM packing."""

    def __init__(self, obj_to_img: Sequence[int], n_images: Optional[int], device):
        ids = list(int(v) for v in obj_to_img)
        O = len(ids)
        self.O = O
        N = (max(ids) + 1) if n_images is None else n_images
        self.N = N

        def dev(lst):
            return torch.tensor(lst, dtype=torch.int32).to(device)

        # crops: boxes grouped by image, ascending box id inside an image
        order = sorted(range(O), key=lambda b: (ids[b], b))
        start = [0] * (N + 1)
        for b in ids:
            start[b + 1] += 1
        for i in range(N):
            start[i + 1] += start[i]
        self.box_to_img = dev(ids)
        self.box_order = dev(order)
        self.img_box_start = dev(start)
        # ConvLSTM sequences: runs of equal ids in batch order (generator_obj_att.py:286-304)
        lens, starts = [], []
        prev = None
        for i, v in enumerate(ids):
            if i == 0 or v != prev:
                starts.append(i)
                lens.append(1)
            else:
                lens[-1] += 1
            prev = v
        self.seq_lens, self.seq_starts = lens, starts
        S = len(lens)
        self.S = S
        perm = sorted(range(S), key=lambda i: (-lens[i], i))
        rank = [0] * S
        for r, i in enumerate(perm):
            rank[i] = r
        T = max(lens) if lens else 0
        n_t = [sum(1 for L in lens if L > t) for t in range(T)]
        offs = [0]
        for n in n_t:
            offs.append(offs[-1] + n)
        self.T, self.n_t, self.offs = T, n_t, offs
        pack_src = [0] * O          # packed row -> object index
        hprev_src = [-1] * O        # packed row -> packed row of the previous step (or -1)
        for t in range(T):
            for j in range(n_t[t]):
                img = perm[j]
                pack_src[offs[t] + j] = starts[img] + t
                if t > 0:
                    hprev_src[offs[t] + j] = offs[t - 1] + j
        unpack_src = [0] * O        # object index -> packed row
        for r, o in enumerate(pack_src):
            unpack_src[o] = r
        final_rows = [offs[lens[i] - 1] + rank[i] for i in range(S)]   # sequence i -> packed row of its last step
        scatter_final = [-1] * O
        for i, r in enumerate(final_rows):
            scatter_final[r] = i
        self.pack_src, self.unpack_src = dev(pack_src), dev(unpack_src)
        self.hprev_src = dev(hprev_src)
        self.final_rows, self.scatter_final = dev(final_rows), dev(scatter_final)


_PLANS: Dict[tuple, BatchPlan] = {}


def get_plan(obj_to_img: torch.Tensor, n_images: Optional[int], device) -> BatchPlan:
    ids = tuple(obj_to_img.tolist())
    key = (ids, n_images, str(device))
    p = _PLANS.get(key)
    if p is None:
        if len(_PLANS) > 64:
            _PLANS.clear()
        p = BatchPlan(ids, n_images, device)
        _PLANS[key] = p
    return p


# ----------------------------------------------------------------------------------------------------------
# box crops
# ----------------------------------------------------------------------------------------------------------
_LINSPACE: Dict[tuple, torch.Tensor] = {}


def crop_weights(S: int, device) -> torch.Tensor:
    """[linspace(1,0,S) | linspace(0,1,S)] built on the CPU in fp32 as bilinear.py:272-275 does, then uploaded."""
    key = (S, str(device))
    w = _LINSPACE.get(key)
    if w is None:
        w = torch.cat([torch.linspace(1, 0, steps=S), torch.linspace(0, 1, steps=S)]).to(device)
        _LINSPACE[key] = w
    return w


class _CropFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feats, boxes, plan: BatchPlan, HH, WW):
        feats = feats.contiguous()
        boxes = boxes.contiguous().float()
        wx, wy = crop_weights(WW, feats.device), crop_weights(HH, feats.device)
        ctx.save_for_backward(boxes, wx, wy)
        ctx.plan, ctx.dims = plan, feats.shape
        return _lib.K.crop_fwd(feats, boxes, plan.box_to_img, wx, wy, HH, WW)

    @staticmethod
    def backward(ctx, dcrops):
        boxes, wx, wy = ctx.saved_tensors
        N, C, H, W = ctx.dims
        p = ctx.plan
        d = _lib.K.crop_bwd(dcrops.contiguous(), boxes, p.img_box_start, p.box_order, wx, wy, N, H, W)
        return d, None, None, None, None


def crop_bbox_batch(feats, bbox, bbox_to_feats, HH, WW=None):
    if WW is None:
        WW = HH
    plan = get_plan(bbox_to_feats.cpu() if bbox_to_feats.is_cuda else bbox_to_feats, feats.shape[0], feats.device)
    return _CropFn.apply(feats, bbox, plan, HH, WW)


# ----------------------------------------------------------------------------------------------------------
# ConvLSTM over per-image object sequences (hoisted input convolution, time-major packing, manual BPTT)
# ----------------------------------------------------------------------------------------------------------
class ConvLSTMLayer:
    def __init__(self, cin, hid, k=5):
        self.cin, self.hid, self.k = cin, hid, k
        self.gx = ConvGeom(cin, 4 * hid, k, k, 1, k // 2, 0, cin + hid)
        self.gh = ConvGeom(hid, 4 * hid, k, k, 1, k // 2, cin, cin + hid)
        self.packs = WeightPacks()


class _ConvLSTMFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, plan: BatchPlan, layers: List[ConvLSTMLayer], *params):
        # params = (w0, b0, w1, b1, ...)
        O, H, W, C0 = x.shape
        P, hw = plan.O, H * W
        saved = []
        xin = _lib.K.permute_rows(x.contiguous().view(O, hw * C0), plan.pack_src, hw * C0).view(P, H, W, C0)
        dev = x.device
        for li, L in enumerate(layers):
            w, b = params[2 * li], params[2 * li + 1]
            hid = L.hid
            pre_x = conv_forward(L.gx, L.packs, w, xin, "cl", "cl", b, None, False)        # (P,H,W,4h)
            h_all = torch.empty((P, H, W, hid), dtype=x.dtype, device=dev)      # activations; cell state and gates fp32
            c_all = torch.empty((P, H, W, hid), dtype=torch.float32, device=dev)
            gates = torch.empty((P, H, W, 4 * hid), dtype=torch.float32, device=dev)
            for t in range(plan.T):
                n, o = plan.n_t[t], plan.offs[t]
                pre_h = c_prev = None
                if t > 0:
                    op = plan.offs[t - 1]
                    pre_h = conv_forward(L.gh, L.packs, w, h_all[op:op + n], "cl", "cl", None, None, False)
                    c_prev = c_all[op:op + n]
                _lib.K.lstm_gates_fwd(pre_x[o:o + n], pre_h, c_prev, n * hw, hid, gates[o:o + n], c_all[o:o + n],
                                      h_all[o:o + n])
            saved.append((xin, h_all, c_all, gates))
            xin = h_all
        hid_last = layers[-1].hid
        out = _lib.K.permute_rows(xin.view(P, hw * hid_last), plan.final_rows, hw * hid_last).view(plan.S, H, W, hid_last)
        ctx.plan, ctx.layers, ctx.saved_acts, ctx.dims = plan, layers, saved, (O, H, W, C0)
        ctx.save_for_backward(*params)
        return out

    @staticmethod
    def backward(ctx, dout):
        plan, layers, saved = ctx.plan, ctx.layers, ctx.saved_acts
        params = ctx.saved_tensors
        O, H, W, C0 = ctx.dims
        P, hw = plan.O, H * W
        hid_last = layers[-1].hid
        dH = _lib.K.permute_rows(dout.contiguous().view(plan.S, hw * hid_last), plan.scatter_final,
                                 hw * hid_last).view(P, H, W, hid_last)
        grads: List[Optional[torch.Tensor]] = [None] * len(params)
        forked, keep = [], []
        for li in range(len(layers) - 1, -1, -1):
            L = layers[li]
            w = params[2 * li]
            hid = L.hid
            xin, h_all, c_all, gates = saved[li]
            dpre = torch.empty((P, H, W, 4 * hid), dtype=dout.dtype, device=dout.device)
            dc_next = None
            n_next = 0
            for t in range(plan.T - 1, -1, -1):
                n, o = plan.n_t[t], plan.offs[t]
                c_prev = c_all[plan.offs[t - 1]:plan.offs[t - 1] + n] if t > 0 else None
                dc_prev = torch.empty((n, H, W, hid), dtype=torch.float32, device=dout.device)
                if n_next > 0:
                    _lib.K.lstm_gates_bwd(dH[o:o + n_next], dc_next, gates[o:o + n_next], c_prev[:n_next] if t > 0 else None,
                                          c_all[o:o + n_next], n_next * hw, hid, dpre[o:o + n_next], dc_prev[:n_next])
                if n > n_next:
                    a, bnd = o + n_next, o + n
                    _lib.K.lstm_gates_bwd(dH[a:bnd], None, gates[a:bnd], c_prev[n_next:n] if t > 0 else None,
                                          c_all[a:bnd], (n - n_next) * hw, hid, dpre[a:bnd], dc_prev[n_next:n])
                if t > 0:
                    op = plan.offs[t - 1]
                    dh_rec = conv_dgrad(L.gh, L.packs, w, dpre[o:o + n], "cl", (H, W), "cl", None, dout.dtype)
                    _lib.K.add(dH[op:op + n], dh_rec, out=dH[op:op + n])
                dc_next, n_next = dc_prev, n
            dpre_op = tc_operand(dpre, _tc_dgrad_ok(L.gx, "cl"))              # one cast for the three GEMMs below
            # the layer's parameter gradients (two weight-gradient GEMMs + the bias column sums) are leaves of this node: they
            # run on a forked stream next to the data gradient and the NEXT layer's latency-bound time-step loop
            side = None
            if SIDE_WGRAD and dout.is_cuda and forks_enabled():
                cur = torch.cuda.current_stream(dout.device)
                side = _phase_streams(dout.device, 4)[3]
                side.wait_stream(cur)
            with (torch.cuda.stream(side) if side is not None else contextlib.nullcontext()):
                gw = torch.empty_like(w)
                conv_wgrad(L.gx, xin, "cl", dpre_op, "cl", gw)
                hprev = _lib.K.permute_rows(h_all.view(P, hw * hid), plan.hprev_src, hw * hid).view(P, H, W, hid)
                conv_wgrad(L.gh, hprev, "cl", dpre_op, "cl", gw)
                grads[2 * li] = gw
                grads[2 * li + 1] = _lib.K.colsum(dpre.view(P * hw, 4 * hid))
            if side is not None:
                forked.append(side)
                keep.append((dpre, dpre_op, xin, hprev, h_all))      # read on the forked stream: alive until the join below
            dxin = conv_dgrad(L.gx, L.packs, w, dpre_op, "cl", (H, W), "cl", None, dout.dtype)
            dH = dxin
        dx = _lib.K.permute_rows(dH.view(P, hw * C0), plan.unpack_src, hw * C0).view(O, H, W, C0)
        for st in forked[-1:]:              # one stream serves every layer (in order): a single join before the node returns
            torch.cuda.current_stream(dout.device).wait_stream(st)
        del keep
        return (dx, None, None) + tuple(grads)


def conv_lstm(x, plan: BatchPlan, layers: List[ConvLSTMLayer], params: Sequence[torch.Tensor]):
    return _ConvLSTMFn.apply(x, plan, layers, *params)


# ----------------------------------------------------------------------------------------------------------
# fused step arithmetic (train64.py:195-252, 284-364): every loss term one launch (value partials + gradient), one launch
# for the total — instead of ~350 elementwise / reduction launches of the PyTorch formulation
# ----------------------------------------------------------------------------------------------------------
LOSS_MAX_BLOCKS = 1024
_LOSS_CONST: Dict[tuple, torch.Tensor] = {}


def _loss_const(values, device) -> torch.Tensor:
    key = (tuple(float(v) for v in values), str(device))
    t = _LOSS_CONST.get(key)
    if t is None:
        t = _LOSS_CONST[key] = torch.tensor(key[0], dtype=torch.float32, device=device)
    return t


class _FusedLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, acc, *inputs):
        terms = _lib.K.loss_total(acc.partials, acc.counts, len(acc.names))
        acc.terms = terms
        ctx.grads = list(acc.grads)
        return terms[len(acc.names)].clone()

    @staticmethod
    def backward(ctx, g):
        grads = ctx.grads
        torch._foreach_mul_(grads, g)          # (the reference calls total.backward(): g == 1; one multi-tensor launch keeps it general)
        return (None,) + tuple(grads)


class FusedLoss:
    """Accumulator of one phase's loss terms.  Each `add_*` call launches one kernel that writes the term's partial sums into
    its slot and returns nothing; `total()` launches the combine kernel and returns the total as an autograd scalar whose
    backward hands every registered input the gradient its kernel already computed.  Term values: `term(name)` after total()."""

    def __init__(self, names, device):
        self.names = list(names)
        n = len(self.names)
        self.partials = torch.empty((n, LOSS_MAX_BLOCKS), dtype=torch.float64, device=device)
        self.counts = torch.zeros((n,), dtype=torch.int32, device=device)
        self.inputs, self.grads, self.terms = [], [], None
        self.device = device

    def _slot(self, name):
        return self.names.index(name)

    def _reg(self, x, grad):
        self.inputs.append(x)
        self.grads.append(grad.view(x.shape))

    @staticmethod
    def _flat(x):
        x = x if x.dtype == torch.float32 else x.float()
        return x.contiguous()

    def add_bce_groups(self, name, x, groups, targets, weights, scale, split_group=None):
        """scale * sum_g weights[g] * mean BCE_with_logits(x[group g], targets[g]); groups >= split_group go to the NEXT name"""
        xc = self._flat(x)
        n = xc.numel() // groups
        g = _lib.K.loss_bce_groups(xc.view(-1), n, groups, groups if split_group is None else split_group,
                                   _loss_const(targets, self.device), _loss_const(weights, self.device), float(scale),
                                   self.partials, self.counts, self._slot(name))
        self._reg(x, g)

    def add_ce_groups(self, name, logits, labels, groups, weights, scale):
        xc = self._flat(logits)
        n = xc.shape[0] // groups
        g = _lib.K.loss_ce_groups(xc, labels.contiguous(), n, groups, _loss_const(weights, self.device), float(scale),
                                  self.partials, self.counts, self._slot(name))
        self._reg(logits, g)

    def add_bce_pos_weight_rows(self, name, logits, targets, sel, n_sel, pos_weight, groups, weights, scale):
        xc = self._flat(logits)
        n = xc.shape[0] // groups
        g = _lib.K.loss_bce_pw_rows(xc, targets.contiguous(), sel, n, groups, int(n_sel), pos_weight,
                                    _loss_const(weights, self.device), float(scale), self.partials, self.counts, self._slot(name))
        self._reg(logits, g)

    def add_l1_rows(self, name, a, b, rows, mask, denom, scale, broadcast_b=False):
        ac = self._flat(a)
        L = ac.numel() // rows
        g = _lib.K.loss_l1_rows(ac, b.contiguous(), rows, L, 0 if broadcast_b else L, mask, float(denom), float(scale),
                                self.partials, self.counts, self._slot(name))
        self._reg(a, g)

    def add_kl(self, name, mu, logvar, scale):
        dmu, dlv = _lib.K.loss_kl(mu.contiguous(), logvar.contiguous(), float(scale), self.partials, self.counts, self._slot(name))
        self._reg(mu, dmu)
        self._reg(logvar, dlv)

    def total(self):
        return _FusedLossFn.apply(self, *self.inputs)

    def term_dict(self, unscale=None):
        """{name: value} (device scalars, views of one tensor); unscale: {name: factor} divides the lambda back out"""
        out = {}
        for i, n in enumerate(self.names):
            v = self.terms[i]
            if unscale and unscale.get(n, 1.0) not in (0.0, 1.0):
                v = v / unscale[n]
            out[n] = v
        return out
