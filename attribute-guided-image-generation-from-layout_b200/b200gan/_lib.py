"""ctypes binding of libb200gan.so (the C ABI declared in include/b200gan.h).

`K` is the kernel namespace every op in this package calls.  There is NO CPU or PyTorch fallback: if the
shared library is missing or a tensor is not on a CUDA device the call raises.  (tests/ may replace `K` with
a CPU emulation of the ABI to validate host-side wiring; nothing in the product does.)
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "csrc", "libb200gan.so")

MODE_PLAIN, MODE_AFFINE, MODE_CBN, MODE_SPADE = 0, 1, 2, 3


class ConvDesc(C.Structure):
    """Mirror of b200_conv_desc (include/b200gan.h)."""
    _fields_ = [
        ("B", C.c_int), ("Qh", C.c_int), ("Qw", C.c_int),
        ("Cin", C.c_int), ("Cout", C.c_int),
        ("Th", C.c_int), ("Tw", C.c_int),
        ("in_sy", C.c_int), ("in_sx", C.c_int),
        ("tap_sy", C.c_int), ("tap_sx", C.c_int),
        ("tap_oy", C.c_int), ("tap_ox", C.c_int),
        ("Hi", C.c_int), ("Wi", C.c_int),
        ("up_shift", C.c_int),
        ("in_sn", C.c_int64), ("in_sh", C.c_int64), ("in_sw", C.c_int64), ("in_sc", C.c_int64),
        ("out_sy", C.c_int), ("out_sx", C.c_int), ("out_oy", C.c_int), ("out_ox", C.c_int),
        ("Ho", C.c_int), ("Wo", C.c_int),
        ("out_sn", C.c_int64), ("out_sh", C.c_int64), ("out_sw", C.c_int64), ("out_sc", C.c_int64),
        ("ldw", C.c_int64),
        ("relu", C.c_int),
        ("scale_rows", C.c_int),
        ("relu_mask", C.c_void_p),
        ("col_stats", C.c_void_p),
        ("col_stats_ld", C.c_int),
    ]


class SNLayer(C.Structure):
    """Mirror of b200_sn_layer (include/b200gan.h)."""
    _fields_ = [("W", C.c_void_p), ("u", C.c_void_p), ("v", C.c_void_p), ("ws", C.c_void_p), ("inv", C.c_void_p),
                ("u_hist", C.c_void_p), ("v_hist", C.c_void_p), ("h", C.c_int), ("w", C.c_int)]


PACK_MT, PACK_CT = 32, 64         # B200_PACK_MT / B200_PACK_CT (include/b200gan.h)


class PackEntry(C.Structure):
    """Mirror of b200_pack_entry (include/b200gan.h)."""
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("ldw", C.c_int64), ("s_m", C.c_int64), ("s_ky", C.c_int64),
                ("s_kx", C.c_int64), ("s_c", C.c_int64), ("dst_bf16", C.c_int32), ("M", C.c_int32), ("Th", C.c_int32),
                ("Tw", C.c_int32), ("C", C.c_int32), ("C_dst", C.c_int32), ("c_off", C.c_int32), ("ky0", C.c_int32),
                ("kx0", C.c_int32), ("kstep", C.c_int32), ("chunk_begin", C.c_int32), ("pad", C.c_int32)]


class B200Error(RuntimeError):
    pass


def _ptr(t: Optional[torch.Tensor]):
    if t is None:
        return None
    if not t.is_cuda:
        raise B200Error("b200gan kernels need CUDA tensors (got a %s tensor); there is no CPU path" % t.device)
    return C.c_void_p(t.data_ptr())


F32, BF16 = 0, 1


def _dt(t: torch.Tensor) -> int:
    """storage type code of an activation tensor (B200_F32 / B200_BF16)"""
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise B200Error("activation tensors are fp32 or bf16, got %s" % t.dtype)


def vec_layout_ok(x2d: torch.Tensor) -> bool:
    """(rows, C) activation layouts the 16-byte vector normalisation kernels take: C/V threads per row, a power of two
    <= 256 (V = 8 bf16 / 4 fp32 channels) — mirrors vec_ok() in csrc/norm.cu"""
    rows, Cc = x2d.shape
    v = 8 if x2d.dtype == torch.bfloat16 else 4
    if Cc % v or rows >= 2 ** 31:
        return False
    tpr = Cc // v
    return 1 <= tpr <= 256 and (tpr & (tpr - 1)) == 0


def _same_dt(*ts):
    ts = [t for t in ts if t is not None]
    d = _dt(ts[0])
    for t in ts[1:]:
        if _dt(t) != d:
            raise B200Error("mixed activation storage types in one call: %s" % [str(t.dtype) for t in ts])
    return d


def _stream():
    """raw handle of torch's current stream on the current device (the fast private accessors: torch.cuda.current_stream()
    costs ~10 us of Python per call, which at ~1700 launches per iteration was a fifth of the eager host time)"""
    return C.c_void_p(torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice()))


class Kernels:
    """Thin, allocation-explicit wrappers: one method per C entry point (or per fixed sequence of them)."""

    def __init__(self, path: str = LIB_PATH):
        if not os.path.exists(path):
            raise B200Error("libb200gan.so not found at %s — build it with `python -c 'import __graft_entry__ as g; "
                            "g.build()'` (no fallback path exists)" % path)
        self.lib = C.CDLL(path)
        self.lib.b200_last_error.restype = C.c_char_p
        self.lib.b200_launch_count.restype = C.c_int64
        self.lib.b200_bn_chunks.argtypes = [C.c_int64, C.c_int]

    # ---- plumbing -------------------------------------------------------------------------------------
    def _check(self, rc: int, name: str):
        if rc != 0:
            raise B200Error("%s failed: %s" % (name, self.lib.b200_last_error().decode()))

    def launch_count(self) -> int:
        return int(self.lib.b200_launch_count())

    def version(self) -> int:
        return int(self.lib.b200_version())

    # ---- box -> layout integer work ------------------------------------------------------------------------
    def rasterize_boxes(self, boxes, H, W):
        if not boxes.is_cuda or boxes.dtype != torch.float32:
            raise B200Error("rasterize_boxes: fp32 CUDA boxes required")
        boxes = boxes.contiguous()
        O_ = boxes.shape[0]
        masks = torch.empty((O_, 1, H, W), dtype=torch.float32, device=boxes.device)
        self._check(self.lib.b200_rasterize_boxes(_ptr(boxes), O_, H, W, _ptr(masks), _stream()), "b200_rasterize_boxes")
        return masks

    def shift_boxes(self, boxes):
        if not boxes.is_cuda or boxes.dtype != torch.float32:
            raise B200Error("shift_boxes: fp32 CUDA boxes required")
        boxes = boxes.contiguous()
        out = torch.empty_like(boxes)
        self._check(self.lib.b200_shift_boxes(_ptr(boxes), boxes.shape[0], _ptr(out), _stream()), "b200_shift_boxes")
        return out

    # ---- data contract ------------------------------------------------------------------------------------
    def one_hot_attributes(self, att_idx, n_att):
        if not att_idx.is_cuda or att_idx.dtype != torch.int64 or att_idx.dim() != 2:
            raise B200Error("one_hot_attributes: (O,A) int64 CUDA tensor required")
        O_, A = att_idx.shape
        out = torch.empty((O_, n_att), dtype=torch.float32, device=att_idx.device)
        self._check(self.lib.b200_one_hot_attributes(_ptr(att_idx), O_, A, int(n_att), _ptr(out), _stream()),
                    "b200_one_hot_attributes")
        return out

    def imagenet_deprocess(self, imgs, inv_std, neg_mean, rescale=True):
        N, Cc, H, W = imgs.shape
        out = torch.empty((N, Cc, H, W), dtype=torch.uint8, device=imgs.device)
        ws = torch.empty((2 * max(N, 1),), dtype=torch.float32, device=imgs.device)
        self._check(self.lib.b200_imagenet_deprocess(_ptr(imgs), N, Cc, H * W, _ptr(inv_std), _ptr(neg_mean),
                                                     int(bool(rescale)), _ptr(out), _ptr(ws), _stream()),
                    "b200_imagenet_deprocess")
        return out

    # ---- masks_to_layout ---------------------------------------------------------------------------------
    def m2l_taps(self, boxes, linx, liny, M):
        O_, H, W = boxes.shape[0], liny.numel(), linx.numel()
        dev = boxes.device
        ix0 = torch.empty((O_, W), dtype=torch.int32, device=dev)
        iy0 = torch.empty((O_, H), dtype=torch.int32, device=dev)
        fx = torch.empty((O_, W), dtype=torch.float32, device=dev)
        fy = torch.empty((O_, H), dtype=torch.float32, device=dev)
        self._check(self.lib.b200_masks_to_layout_taps(_ptr(boxes), _ptr(linx), _ptr(liny), _ptr(ix0), _ptr(iy0), _ptr(fx),
                                                       _ptr(fy), int(M), H, W, O_, _stream()), "b200_masks_to_layout_taps")
        return ix0, iy0, fx, fy

    def m2l_fwd(self, vecs, boxes, masks, img_obj_start, obj_order, linx, liny, N):
        O_, D = vecs.shape
        M, H, W = masks.shape[-1], liny.numel(), linx.numel()
        for t in (vecs, boxes, masks):
            if t.dtype != torch.float32 or not t.is_contiguous():
                raise B200Error("masks_to_layout: contiguous fp32 tensors required")
        out = torch.empty((N, D, H, W), dtype=torch.float32, device=vecs.device)
        self._check(self.lib.b200_masks_to_layout_fwd(_ptr(vecs), _ptr(boxes), _ptr(masks), _ptr(img_obj_start),
                                                      _ptr(obj_order), _ptr(linx), _ptr(liny), _ptr(out), int(N), O_, D, M,
                                                      H, W, _stream()), "b200_masks_to_layout_fwd")
        return out

    def m2l_bwd(self, dout, vecs, boxes, masks, obj_to_img, linx, liny, need_vecs=True, need_masks=True):
        O_, D = vecs.shape
        N, _, H, W = dout.shape
        M = masks.shape[-1]
        dev = vecs.device
        dvecs = torch.empty((O_, D), dtype=torch.float32, device=dev) if need_vecs else None
        dmasks = torch.empty((O_, M, M), dtype=torch.float32, device=dev) if need_masks else None
        ws = torch.empty((O_ * H * W,), dtype=torch.float32, device=dev) if need_masks else None
        self._check(self.lib.b200_masks_to_layout_bwd(_ptr(dout), _ptr(vecs), _ptr(boxes), _ptr(masks), _ptr(obj_to_img),
                                                      _ptr(linx), _ptr(liny), _ptr(dvecs), _ptr(dmasks), _ptr(ws), N, O_, D,
                                                      M, H, W, _stream()), "b200_masks_to_layout_bwd")
        return dvecs, dmasks

    # ---- crops ----------------------------------------------------------------------------------------
    def crop_fwd(self, feats, boxes, box_to_img, wx, wy, HH, WW):
        N, Cc, H, W = feats.shape
        B = boxes.shape[0]
        out = torch.empty((B, Cc, HH, WW), dtype=torch.float32, device=feats.device)
        self._check(self.lib.b200_crop_fwd(_ptr(feats), _ptr(boxes), _ptr(box_to_img), _ptr(wx), _ptr(wy), _ptr(out),
                                           N, Cc, H, W, B, HH, WW, _stream()), "b200_crop_fwd")
        return out

    def crop_set_staged(self, enable: bool) -> bool:
        """shared-memory staged crop kernels (default) vs the element-wise ones; returns the previous setting"""
        return bool(self.lib.b200_crop_set_staged(int(bool(enable))))

    def crop_taps(self, boxes, wx, wy, H, W, HH, WW):
        B = boxes.shape[0]
        dev = boxes.device
        ix0 = torch.empty((B, WW), dtype=torch.int32, device=dev)
        iy0 = torch.empty((B, HH), dtype=torch.int32, device=dev)
        fx = torch.empty((B, WW), dtype=torch.float32, device=dev)
        fy = torch.empty((B, HH), dtype=torch.float32, device=dev)
        self._check(self.lib.b200_crop_taps(_ptr(boxes), _ptr(wx), _ptr(wy), _ptr(ix0), _ptr(iy0), _ptr(fx), _ptr(fy),
                                            H, W, B, HH, WW, _stream()), "b200_crop_taps")
        return ix0, iy0, fx, fy

    def crop_bwd(self, dcrops, boxes, img_box_start, box_order, wx, wy, N, H, W):
        B, Cc, HH, WW = dcrops.shape
        dfeats = torch.empty((N, Cc, H, W), dtype=torch.float32, device=dcrops.device)
        ws = torch.empty(((B * Cc * H * WW + 3) // 4 * 4 + 4 * max(1, B),), dtype=torch.float32, device=dcrops.device)
        self._check(self.lib.b200_crop_bwd(_ptr(dcrops), _ptr(boxes), _ptr(img_box_start), _ptr(box_order), _ptr(wx),
                                           _ptr(wy), _ptr(dfeats), _ptr(ws), N, Cc, H, W, B, HH, WW, _stream()),
                    "b200_crop_bwd")
        return dfeats

    # ---- gather-GEMM convolutions -----------------------------------------------------------------------
    def cast_bf16(self, x):
        """fp32 -> bf16 copy of a contiguous tensor (the operand format of the tcgen05 gather-GEMMs)"""
        y = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
        self._check(self.lib.b200_cast_bf16(_ptr(x), _ptr(y), C.c_int64(x.numel()), _stream()), "b200_cast_bf16")
        return y

    def split_bf16(self, x):
        """(hi, lo) bf16 tensors with hi + lo ~ x to 16 mantissa bits"""
        hi = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
        lo = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
        self._check(self.lib.b200_split_bf16(_ptr(x), _ptr(hi), _ptr(lo), C.c_int64(x.numel()), _stream()), "b200_split_bf16")
        return hi, lo

    def im2col_pack(self, x, x_strides, N, Hx, Wx, Cx, kh, kw, stride, pad, Hy, Wy, Kp, flip=False):
        """bf16 [N*Hy*Wy, Kp] im2col matrix of a few-channel input (column = (ky*kw+kx)*Cx + c, zero padded to Kp);
        flip: the transposed window walk (stride 1) over an output gradient"""
        out = torch.empty((N * Hy * Wy, Kp), dtype=torch.bfloat16, device=x.device)
        sn, sh, sw, sc = x_strides
        self._check(self.lib.b200_im2col_pack(_ptr(x), _dt(x), C.c_int64(N), Hx, Wx, Cx, C.c_int64(sn), C.c_int64(sh),
                                              C.c_int64(sw), C.c_int64(sc), kh, kw, stride, pad, Hy, Wy, Kp,
                                              int(bool(flip)), _ptr(out), _stream()), "b200_im2col_pack")
        return out

    def rowsum(self, x2d):
        rows, L = x2d.shape
        out = torch.empty((rows,), dtype=torch.float32, device=x2d.device)
        self._check(self.lib.b200_rowsum(_ptr(x2d), _dt(x2d), C.c_int64(rows), C.c_int64(L), _ptr(out), _stream()),
                    "b200_rowsum")
        return out

    def conv_stats_ok(self, desc: ConvDesc, tc) -> bool:
        """will this launch run on the persistent tcgen05 kernel (whose epilogue can accumulate batch-norm statistics)?"""
        tc = int(tc)
        return bool(tc) and bool(self.lib.b200_conv_tc_stats_ok(C.byref(desc), 4 if tc == 2 else 2))

    def conv_stats_buffer(self, desc: ConvDesc, device):
        """(slabs, ld, 2) fp32 buffer for b200_conv_desc.col_stats: one (sum, sum of squares) pair per 32-row slab and channel"""
        M = desc.B * desc.Qh * desc.Qw
        nt = self.conv_tc_ntile(desc.Cout)
        ld = (desc.Cout + nt - 1) // nt * nt
        return torch.empty(((M + 127) // 128 * 4, ld, 2), dtype=torch.float32, device=device)

    def bn_stats_slabs(self, col_stats, rows, Cc, running_mean, running_var, momentum, groups=1):
        dev = col_stats.device
        mean = torch.empty((groups, Cc), dtype=torch.float32, device=dev)
        var = torch.empty((groups, Cc), dtype=torch.float32, device=dev)
        self._check(self.lib.b200_bn_stats_slabs(_ptr(col_stats), C.c_int64(col_stats.shape[0]), int(col_stats.shape[1]),
                                                 C.c_int64(rows), int(Cc), int(groups), _ptr(mean), _ptr(var), _ptr(running_mean),
                                                 _ptr(running_var), C.c_float(momentum), _stream()), "b200_bn_stats_slabs")
        return mean, var

    def conv_gemm(self, desc: ConvDesc, inp, wmat, bias, scale, out, tc, mask=None, stats=None):
        """tc: 0 = fp32 CUDA-core kernel, 1 = tcgen05 bf16 operands, 2 = tcgen05 tf32 (fp32 tensors; falls back to the CUDA-core
        kernel for gathers the TMA im2col mode cannot express).  mask (tcgen05 paths only): tensor laid out like `out`;
        outputs are zeroed where it is not > 0"""
        tc = int(tc)
        if mask is not None:
            if not tc or mask.dtype != out.dtype or mask.shape != out.shape or not mask.is_contiguous():
                raise B200Error("conv_gemm: relu mask must match the output (tcgen05 path)")
            desc.relu_mask = mask.data_ptr()
        else:
            desc.relu_mask = None
        if stats is not None:            # (the caller checked conv_stats_ok: persistent tcgen05 kernel)
            desc.col_stats, desc.col_stats_ld = stats.data_ptr(), int(stats.shape[1])
        else:
            desc.col_stats, desc.col_stats_ld = None, 0
        if tc == 2:
            if inp.dtype != torch.float32 or wmat.dtype != torch.float32 or out.dtype != torch.float32:
                raise B200Error("conv_gemm(tf32): fp32 activation, weight matrix and output required")
            if not self.lib.b200_conv_tf32_ok(C.byref(desc), 0):
                if mask is not None:
                    raise B200Error("conv_gemm(tf32): the fused ReLU mask needs the tcgen05 path")
                tc = 0
            else:
                splits = int(self.lib.b200_conv_tf32_splits(C.byref(desc)))
                ws = None
                if splits > 1:
                    nt = self.conv_tc_ntile(desc.Cout)
                    ldo = (desc.Cout + nt - 1) // nt * nt
                    ws = torch.empty((splits * desc.B * desc.Qh * desc.Qw * ldo,), dtype=torch.float32, device=inp.device)
                self._check(self.lib.b200_conv_gemm_tf32(C.byref(desc), _ptr(inp), _ptr(wmat), _ptr(bias), _ptr(scale),
                                                         _ptr(out), _ptr(ws), splits, _stream()), "b200_conv_gemm_tf32")
                return
        if not tc:
            self._check(self.lib.b200_conv_gemm_f32(C.byref(desc), _ptr(inp), _dt(inp), _ptr(wmat), _ptr(bias), _ptr(scale),
                                                    _ptr(out), _dt(out), _stream()), "b200_conv_gemm_f32")
            return
        if inp.dtype != torch.bfloat16:
            raise B200Error("conv_gemm(tc): the activation operand must be bf16 (use cast_bf16)")
        splits = int(self.lib.b200_conv_tc_splits(C.byref(desc)))
        ws = None
        if splits > 1:
            nt = self.conv_tc_ntile(desc.Cout)
            ldo = (desc.Cout + nt - 1) // nt * nt
            ws = torch.empty((splits * desc.B * desc.Qh * desc.Qw * ldo,), dtype=torch.float32, device=inp.device)
        self._check(self.lib.b200_conv_gemm_tc(C.byref(desc), _ptr(inp), _ptr(wmat), _ptr(bias), _ptr(scale), _ptr(out),
                                               int(out.dtype == torch.bfloat16), _ptr(ws), splits, _stream()),
                    "b200_conv_gemm_tc")

    def conv_tc_set_im2col(self, enable: bool) -> bool:
        return bool(self.lib.b200_conv_tc_set_im2col(int(bool(enable))))

    def conv_tc_set_persistent(self, mode) -> int:
        """0 off, 1 one CTA per SM, 2 two CTAs per SM, 3 two CTAs per SM for short K loops; returns the previous mode"""
        return int(self.lib.b200_conv_tc_set_persistent(int(mode)))

    def conv_tc_set_halo(self, enable: bool) -> bool:
        return bool(self.lib.b200_conv_tc_set_halo(int(bool(enable))))

    def conv_tc_ntile(self, cout: int) -> int:
        return int(self.lib.b200_conv_tc_ntile(int(cout)))

    def wgrad_gemm(self, desc: ConvDesc, P, G, ws, splits: int, tc):
        tc = int(tc)
        if tc == 1 and (P.dtype != torch.bfloat16 or G.dtype != torch.bfloat16):
            raise B200Error("wgrad_gemm(tc): both operands must be bf16 (use cast_bf16)")
        if tc == 2:
            if P.dtype != torch.float32 or G.dtype != torch.float32:
                raise B200Error("wgrad_gemm(tf32): fp32 operands required")
            if self.lib.b200_conv_tf32_ok(C.byref(desc), 1):
                self._check(self.lib.b200_wgrad_gemm_tf32(C.byref(desc), _ptr(P), _ptr(G), _ptr(ws), int(splits), _stream()),
                            "b200_wgrad_gemm_tf32")
                return
            tc = 0
        if tc:
            self._check(self.lib.b200_wgrad_gemm_tc(C.byref(desc), _ptr(P), _ptr(G), _ptr(ws), int(splits), _stream()),
                        "b200_wgrad_gemm_tc")
        else:
            self._check(self.lib.b200_wgrad_gemm_f32(C.byref(desc), _ptr(P), _dt(P), _ptr(G), _dt(G), _ptr(ws), int(splits),
                                                     _stream()), "b200_wgrad_gemm_f32")

    def wgrad_reduce(self, ws, splits, M, Th, Tw, Cc, dst, dst_offset, s_m, s_ty, s_tx, s_c, scale=None,
                     accumulate=False, ws_row_offset=0, ws_rows=None):
        """dst (+dst_offset elements) receives rows [ws_row_offset, ws_row_offset+M) of the partial results;
        ws holds `ws_rows` rows per split (default M)."""
        rows_total = M if ws_rows is None else ws_rows
        K = Th * Tw * Cc
        if not (dst.is_cuda and ws.is_cuda):
            raise B200Error("wgrad_reduce: CUDA tensors required")
        wptr = C.c_void_p(ws.data_ptr() + 4 * int(ws_row_offset) * K)
        dptr = C.c_void_p(dst.data_ptr() + 4 * int(dst_offset))
        self._check(self.lib.b200_wgrad_reduce(wptr, int(splits), C.c_int64(rows_total * K), int(M), int(Th), int(Tw),
                                               int(Cc), dptr, C.c_int64(s_m), C.c_int64(s_ty), C.c_int64(s_tx),
                                               C.c_int64(s_c), _ptr(scale), int(bool(accumulate)), _stream()),
                    "b200_wgrad_reduce")

    def pack_weight(self, src, src_offset, dst, dst_row_offset, bf16, M, Mpad, Th, Tw, Cc, ldw, s_m, s_ky, s_kx, s_c,
                    ky0=0, kx0=0, kstep=1, C_dst=0, c_off=0):
        esz = 2 if bf16 else 4
        sptr = C.c_void_p(src.data_ptr() + 4 * int(src_offset))
        dptr = C.c_void_p(dst.data_ptr() + esz * int(dst_row_offset) * int(ldw))
        if not (src.is_cuda and dst.is_cuda):
            raise B200Error("pack_weight: CUDA tensors required")
        self._check(self.lib.b200_pack_weight(sptr, dptr, int(bool(bf16)), int(M), int(Mpad), int(Th), int(Tw), int(Cc),
                                              C.c_int64(ldw), C.c_int64(s_m), C.c_int64(s_ky), C.c_int64(s_kx),
                                              C.c_int64(s_c), int(ky0), int(kx0), int(kstep), int(C_dst), int(c_off),
                                              _stream()), "b200_pack_weight")

    def pack_table(self, recipes, device):
        """(pinned host image, device copy, n_entries, total_chunks) of the b200_pack_entry table for `recipes` (objects
        with the attributes of ops.PackRecipe); the copy is an asynchronous pinned copy (capturable)"""
        arr = (PackEntry * len(recipes))()
        chunks = 0
        owner = []
        for i, r in enumerate(recipes):
            esz = 2 if r.bf16 else 4
            cd = r.C_dst if r.C_dst > 0 else r.C
            arr[i] = PackEntry(r.src.data_ptr() + 4 * int(r.src_offset),
                               r.dst.data_ptr() + esz * int(r.dst_row_offset) * int(r.ldw), r.ldw, r.s_m, r.s_ky, r.s_kx,
                               r.s_c, int(bool(r.bf16)), r.M, r.Th, r.Tw, r.C, cd, r.c_off if r.C_dst > 0 else 0, r.ky0,
                               r.kx0, r.kstep, chunks, 0)
            n = -(-r.M // PACK_MT) * -(-r.C // PACK_CT)
            owner.append(torch.full((n,), i, dtype=torch.int32))
            chunks += n
        nbytes = C.sizeof(arr)
        host = torch.empty((nbytes,), dtype=torch.uint8).pin_memory()
        C.memmove(host.data_ptr(), C.addressof(arr), nbytes)
        dev = torch.empty((nbytes,), dtype=torch.uint8, device=device)
        dev.copy_(host, non_blocking=True)
        owner_host = torch.cat(owner).pin_memory()
        owner_dev = torch.empty_like(owner_host, device=device)
        owner_dev.copy_(owner_host, non_blocking=True)
        return (host, owner_host), (dev, owner_dev), len(recipes), chunks

    def pack_weight_multi(self, table_dev, n_entries, total_chunks):
        dev, owner = table_dev
        self._check(self.lib.b200_pack_weight_multi(_ptr(dev), int(n_entries), int(total_chunks), _ptr(owner), _stream()),
                    "b200_pack_weight_multi")

    # ---- normalisation --------------------------------------------------------------------------------------
    def bn_chunks(self, rows, Cc):
        return int(self.lib.b200_bn_chunks(C.c_int64(rows), int(Cc)))

    def bn_stats(self, x2d, running_mean, running_var, momentum, groups=1):
        rows, Cc = x2d.shape
        dev = x2d.device
        mean = torch.empty((groups, Cc), dtype=torch.float32, device=dev)
        var = torch.empty((groups, Cc), dtype=torch.float32, device=dev)
        ws = torch.empty((2 * Cc * groups * self.bn_chunks(rows // groups, Cc),), dtype=torch.float64, device=dev)
        self._check(self.lib.b200_bn_stats(_ptr(x2d), _dt(x2d), C.c_int64(rows), Cc, int(groups), _ptr(mean), _ptr(var),
                                           _ptr(running_mean), _ptr(running_var), C.c_float(momentum), _ptr(ws),
                                           _stream()), "b200_bn_stats")
        return mean, var

    def norm_fwd(self, x2d, mean, var, eps, mode, gamma, beta, idx, rows_per_seg, residual, relu, groups=1):
        rows, Cc = x2d.shape
        y = torch.empty_like(x2d)
        dt = _same_dt(x2d, residual, gamma if mode == MODE_SPADE else None)
        self._check(self.lib.b200_norm_fwd(_ptr(x2d), _ptr(y), dt, C.c_int64(rows), Cc, int(groups), _ptr(mean), _ptr(var),
                                           C.c_float(eps), int(mode), _ptr(gamma), _ptr(beta), _ptr(idx),
                                           int(rows_per_seg), _ptr(residual), int(bool(relu)), _stream()),
                    "b200_norm_fwd")
        return y

    def norm_bwd(self, dy, x2d, y, mean, var, eps, mode, gamma, idx, rows_per_seg, relu, num_classes, groups=1, sync=None):
        """reduce -> finalize -> apply.  Returns dx, dgamma, dbeta, dtable, dgb (None where not applicable).
        sync: optional callable applied to the (groups*C*2,) fp32 tensor of per-group (sum dxhat, sum dxhat*xhat) between
        finalize and apply (synchronised batch norm: all-reduce over the data-parallel ranks)."""
        rows, Cc = x2d.shape
        dev = x2d.device
        rpg = rows // groups
        dt = _same_dt(dy, x2d, y, gamma if mode == MODE_SPADE else None)
        if mode == MODE_CBN:
            seg = int(rows_per_seg)
        else:
            nchunks = self.bn_chunks(rpg, Cc)
            seg = (rpg + nchunks - 1) // nchunks
        nseg = groups * ((rpg + seg - 1) // seg)
        seg_sums = torch.empty((nseg * Cc * 2,), dtype=torch.float64, device=dev)
        self._check(self.lib.b200_norm_bwd_reduce(_ptr(dy), _ptr(x2d), _ptr(y), dt, C.c_int64(rows), Cc, int(groups),
                                                  _ptr(mean), _ptr(var), C.c_float(eps), int(mode), _ptr(gamma),
                                                  _ptr(idx), seg, int(relu), _ptr(seg_sums), _stream()),
                    "b200_norm_bwd_reduce")
        s = torch.empty((groups * Cc * 2,), dtype=torch.float32, device=dev)
        dgamma = dbeta = dtable = dgb = None
        if mode == MODE_AFFINE:
            dgamma = torch.empty((Cc,), dtype=torch.float32, device=dev)
            dbeta = torch.empty((Cc,), dtype=torch.float32, device=dev)
        if mode == MODE_CBN:
            dtable = torch.empty((num_classes, 2 * Cc), dtype=torch.float32, device=dev)
        self._check(self.lib.b200_norm_bwd_finalize(_ptr(seg_sums), nseg, Cc, int(groups), int(mode), _ptr(gamma),
                                                    _ptr(idx), int(num_classes), _ptr(s), _ptr(dgamma), _ptr(dbeta),
                                                    _ptr(dtable), _stream()), "b200_norm_bwd_finalize")
        if sync is not None:
            s = sync(s)
        dx = torch.empty_like(x2d)
        if mode == MODE_SPADE:
            dgb = torch.empty((rows, 2 * Cc), dtype=x2d.dtype, device=dev)
        self._check(self.lib.b200_norm_bwd_apply(_ptr(dy), _ptr(x2d), _ptr(y), _ptr(dx), dt, C.c_int64(rows), Cc,
                                                 int(groups), _ptr(mean), _ptr(var), C.c_float(eps), int(mode),
                                                 _ptr(gamma), _ptr(idx), int(rows_per_seg) if mode == MODE_CBN else 1,
                                                 int(relu), _ptr(s), _ptr(dgb), _stream()), "b200_norm_bwd_apply")
        return dx, dgamma, dbeta, dtable, dgb

    # ---- elementwise / pooling / layout ------------------------------------------------------------------------
    def relu_fwd(self, x):
        y = torch.empty_like(x)
        self._check(self.lib.b200_relu_fwd(_ptr(x), _ptr(y), C.c_int64(x.numel()), _dt(x), _stream()), "b200_relu_fwd")
        return y

    def relu_bwd(self, dy, y):
        dx = torch.empty_like(dy)
        self._check(self.lib.b200_relu_bwd(_ptr(dy), _ptr(y), _ptr(dx), C.c_int64(dy.numel()), _same_dt(dy, y), _stream()),
                    "b200_relu_bwd")
        return dx

    def add(self, a, b, out=None):
        if out is None:
            out = torch.empty_like(a)
        self._check(self.lib.b200_add(_ptr(a), _ptr(b), _ptr(out), C.c_int64(a.numel()), _same_dt(a, b, out), _stream()),
                    "b200_add")
        return out

    def pool_fwd(self, x, N, H, W, Cc, f, scale):
        y = torch.empty((N, H // f, W // f, Cc), dtype=x.dtype, device=x.device)
        self._check(self.lib.b200_pool_fwd(_ptr(x), _ptr(y), N, H, W, Cc, f, C.c_float(scale), _dt(x), _stream()),
                    "b200_pool_fwd")
        return y

    def pool_add_fwd(self, a, b, N, H, W, Cc, f, scale, relu=False):
        _same_dt(a, b)
        y = torch.empty((N, H // f, W // f, Cc), dtype=a.dtype, device=a.device)
        self._check(self.lib.b200_pool_add_fwd(_ptr(a), _ptr(b), _ptr(y), N, H, W, Cc, f, C.c_float(scale), int(bool(relu)),
                                               _dt(a), _stream()), "b200_pool_add_fwd")
        return y

    def add_relu(self, a, b):
        dt = _same_dt(a, b)
        out = torch.empty_like(a)
        self._check(self.lib.b200_add_relu(_ptr(a), _ptr(b), _ptr(out), C.c_int64(a.numel()), dt, _stream()), "b200_add_relu")
        return out

    def unpool_masked_fwd(self, x, mask, N, H, W, Cc, f, scale):
        _same_dt(x, mask)
        y = torch.empty((N, H * f, W * f, Cc), dtype=x.dtype, device=x.device)
        self._check(self.lib.b200_unpool_masked_fwd(_ptr(x), _ptr(mask), _ptr(y), N, H, W, Cc, f, C.c_float(scale), _dt(x),
                                                    _stream()), "b200_unpool_masked_fwd")
        return y

    def unpool_fwd(self, x, N, H, W, Cc, f, scale):
        y = torch.empty((N, H * f, W * f, Cc), dtype=x.dtype, device=x.device)
        self._check(self.lib.b200_unpool_fwd(_ptr(x), _ptr(y), N, H, W, Cc, f, C.c_float(scale), _dt(x), _stream()),
                    "b200_unpool_fwd")
        return y

    def concat_fwd(self, a, Ca, a_div, b, Cb, b_div, rows):
        out = torch.empty((rows, Ca + Cb), dtype=a.dtype, device=a.device)
        self._check(self.lib.b200_concat_fwd(_ptr(a), Ca, a_div, _ptr(b), Cb, b_div, _ptr(out), C.c_int64(rows),
                                             _same_dt(a, b), _stream()),
                    "b200_concat_fwd")
        return out

    def concat_bwd(self, dout, Ca, a_div, Cb, b_div, rows, need_a=True, need_b=True):
        dev = dout.device
        da = torch.empty((rows // a_div, Ca), dtype=dout.dtype, device=dev) if need_a else None
        db = torch.empty((rows // b_div, Cb), dtype=dout.dtype, device=dev) if need_b else None
        self._check(self.lib.b200_concat_bwd(_ptr(dout), Ca, a_div, _ptr(da), Cb, b_div, _ptr(db), C.c_int64(rows),
                                             _dt(dout), _stream()), "b200_concat_bwd")
        return da, db

    def gather_rows(self, table, idx):
        rows, D = idx.shape[0], table.shape[1]
        out = torch.empty((rows, D), dtype=torch.float32, device=table.device)
        self._check(self.lib.b200_gather_rows(_ptr(table), _ptr(idx), _ptr(out), rows, D, _stream()), "b200_gather_rows")
        return out

    def scatter_rows(self, dout, idx, num_classes):
        rows, D = dout.shape
        dtable = torch.empty((num_classes, D), dtype=torch.float32, device=dout.device)
        self._check(self.lib.b200_scatter_rows(_ptr(dout), _ptr(idx), _ptr(dtable), rows, D, num_classes, _stream()),
                    "b200_scatter_rows")
        return dtable

    def permute_rows(self, x, src_row, rowlen):
        rows = src_row.shape[0]
        out = torch.empty((rows, rowlen), dtype=x.dtype, device=x.device)
        self._check(self.lib.b200_permute_rows(_ptr(x), _ptr(src_row), _ptr(out), rows,
                                               C.c_int64(rowlen * x.element_size()), _stream()), "b200_permute_rows")
        return out

    def mask_outer_fwd(self, v, mask, O, H, W, Cc, out_dtype=torch.float32):
        out = torch.empty((O, H + 2, W + 2, Cc), dtype=out_dtype, device=v.device)
        self._check(self.lib.b200_mask_outer_fwd(_ptr(v), _ptr(mask), _ptr(out), O, H, W, Cc, _dt(out), _stream()),
                    "b200_mask_outer_fwd")
        return out

    def mask_outer_bwd(self, dout, mask, O, H, W, Cc):
        dv = torch.empty((O, Cc), dtype=torch.float32, device=dout.device)
        self._check(self.lib.b200_mask_outer_bwd(_ptr(dout), _ptr(mask), _ptr(dv), O, H, W, Cc, _dt(dout), _stream()),
                    "b200_mask_outer_bwd")
        return dv

    def lstm_gates_fwd(self, pre_x, pre_h, c_prev, rows, hid, gates=None, c_out=None, h_out=None):
        dev = pre_x.device
        if gates is None:
            gates = torch.empty((rows, 4 * hid), dtype=torch.float32, device=dev)
            c_out = torch.empty((rows, hid), dtype=torch.float32, device=dev)
            h_out = torch.empty((rows, hid), dtype=pre_x.dtype, device=dev)
        self._check(self.lib.b200_lstm_gates_fwd(_ptr(pre_x), _ptr(pre_h), _ptr(c_prev), _ptr(gates), _ptr(c_out),
                                                 _ptr(h_out), C.c_int64(rows), hid, _same_dt(pre_x, pre_h, h_out),
                                                 _stream()), "b200_lstm_gates_fwd")
        return gates, c_out, h_out

    def lstm_gates_bwd(self, dh, dc_next, gates, c_prev, c_out, rows, hid, dpre=None, dc_prev=None):
        dev = dh.device
        if dpre is None:
            dpre = torch.empty((rows, 4 * hid), dtype=dh.dtype, device=dev)
        if dc_prev is None:
            dc_prev = torch.empty((rows, hid), dtype=torch.float32, device=dev)
        self._check(self.lib.b200_lstm_gates_bwd(_ptr(dh), _ptr(dc_next), _ptr(gates), _ptr(c_prev), _ptr(c_out),
                                                 _ptr(dpre), _ptr(dc_prev), C.c_int64(rows), hid, _same_dt(dh, dpre),
                                                 _stream()),
                    "b200_lstm_gates_bwd")
        return dpre, dc_prev

    def reparam_fwd(self, mu, logvar, eps):
        z = torch.empty_like(mu)
        self._check(self.lib.b200_reparam_fwd(_ptr(mu), _ptr(logvar), _ptr(eps), _ptr(z), C.c_int64(mu.numel()), _stream()),
                    "b200_reparam_fwd")
        return z

    def reparam_bwd(self, dz, logvar, eps):
        dmu = torch.empty_like(dz)
        dlv = torch.empty_like(dz)
        self._check(self.lib.b200_reparam_bwd(_ptr(dz), _ptr(logvar), _ptr(eps), _ptr(dmu), _ptr(dlv),
                                              C.c_int64(dz.numel()), _stream()), "b200_reparam_bwd")
        return dmu, dlv

    def transpose(self, x, B, R, Cc):
        y = torch.empty((B, Cc, R), dtype=torch.float32, device=x.device)
        self._check(self.lib.b200_transpose(_ptr(x), _ptr(y), B, R, Cc, _stream()), "b200_transpose")
        return y

    def colsum(self, x2d):
        rows, Cc = x2d.shape
        out = torch.empty((Cc,), dtype=torch.float32, device=x2d.device)
        ws = torch.empty((Cc * self.bn_chunks(rows, Cc),), dtype=torch.float64, device=x2d.device)
        self._check(self.lib.b200_colsum(_ptr(x2d), C.c_int64(rows), Cc, _dt(x2d), _ptr(out), _ptr(ws), _stream()),
                    "b200_colsum")
        return out

    # ---- spectral norm --------------------------------------------------------------------------------------
    def sn_power_iter(self, W2d_param, h, w, u, v, do_iter, eps, inv_out=None, sigma_out=None):
        """one power iteration (in place on u, v); writes 1/sigma to inv_out[0] (a 1-element view) and returns it"""
        dev = u.device
        if inv_out is None:
            inv_out = torch.empty((1,), dtype=torch.float32, device=dev)
        ws = torch.empty((8 * w + h,), dtype=torch.float32, device=dev)
        self._check(self.lib.b200_sn_power_iter(_ptr(W2d_param), h, w, _ptr(u), _ptr(v), int(bool(do_iter)),
                                                C.c_float(eps), _ptr(sigma_out), _ptr(inv_out), _ptr(ws), _stream()),
                    "b200_sn_power_iter")
        return inv_out

    def sn_grad(self, g, W, u, v, inv_sigma, h, w, dW=None, accumulate=False):
        if dW is None:
            dW = torch.empty_like(W)
        ws = torch.empty((1024,), dtype=torch.float64, device=W.device)
        self._check(self.lib.b200_sn_grad(_ptr(g), _ptr(W), _ptr(u), _ptr(v), _ptr(inv_sigma), _ptr(dW), h, w,
                                          int(bool(accumulate)), _ptr(ws), _stream()), "b200_sn_grad")
        return dW

    def sn_wgrad_finish(self, ws, groups, spg, Cy, T, Cx, W, u_hist, v_hist, inv, dW, pooled_taps=None):
        """dW of `groups` batched calls of a spectral-normalised layer from the group-aligned partials of one wgrad launch.
        pooled_taps = (Th, Tw) of the PARAMETER: `ws` are the partials of the folded (Th+1) x (Tw+1) stride-2 convolution
        (fold_pool_weight) and T is ignored"""
        dev = W.device
        parts = int(self.lib.b200_sn_wgrad_parts(int(Cy), int(Cx)))
        dot = torch.empty((groups * parts,), dtype=torch.float64, device=dev)
        if pooled_taps is not None:
            th, tw = pooled_taps
            self._check(self.lib.b200_sn_wgrad_finish_pooled(_ptr(ws), int(groups), int(spg),
                                                             C.c_int64(Cy * (th + 1) * (tw + 1) * Cx), int(Cy), int(th), int(tw),
                                                             int(Cx), _ptr(W), _ptr(u_hist), _ptr(v_hist), _ptr(inv),
                                                             _ptr(dot), _ptr(dW), _stream()), "b200_sn_wgrad_finish_pooled")
            return dW
        self._check(self.lib.b200_sn_wgrad_finish(_ptr(ws), int(groups), int(spg), C.c_int64(Cy * T * Cx), int(Cy), int(T),
                                                  int(Cx), _ptr(W), _ptr(u_hist), _ptr(v_hist), _ptr(inv), None,
                                                  _ptr(dot), _ptr(dW), _stream()), "b200_sn_wgrad_finish")
        return dW

    def fold_pool_weight(self, w, w4):
        """w (.., kh, kw) fp32 contiguous -> w4 (.., kh+1, kw+1): the weight of the stride-2 convolution equal to
        avg_pool2(conv(x; w)) (see b200gan.h)"""
        kh, kw = int(w.shape[-2]), int(w.shape[-1])
        if not (w.is_cuda and w4.is_cuda and w.dtype == torch.float32 and w4.dtype == torch.float32
                and w.is_contiguous() and w4.is_contiguous()):
            raise B200Error("fold_pool_weight: contiguous fp32 CUDA tensors required")
        filters = w.numel() // (kh * kw)
        if w4.numel() != filters * (kh + 1) * (kw + 1):
            raise B200Error("fold_pool_weight: bad destination shape")
        self._check(self.lib.b200_fold_pool_weight(_ptr(w), _ptr(w4), C.c_int64(filters), kh, kw, _stream()),
                    "b200_fold_pool_weight")
        return w4

    def sn_table(self, layers, stage, ws, iters):
        """device array of b200_sn_layer for `layers` = [(W, u, v, h, w, stage_off, ws_off)]; per layer the stage buffer
        holds [inv (iters) | u_hist (iters*h) | v_hist (iters*w)] at stage_off (floats)"""
        arr = (SNLayer * len(layers))()
        for i, (W, u, v, h, w, so, wo) in enumerate(layers):
            base = stage.data_ptr() + 4 * so
            arr[i] = SNLayer(W.data_ptr(), u.data_ptr(), v.data_ptr(), ws.data_ptr() + 4 * wo, base, base + 4 * iters,
                             base + 4 * (iters + iters * h), h, w)
        host = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
        return host.to(stage.device)

    def sn_power_iter_multi(self, table, layers, stage, ws, iters, do_iter, eps):
        """`iters` power iterations of every layer in `table` (built by sn_table from `layers`), 4 launches each"""
        max_h = max(l[3] for l in layers)
        max_w = max(l[4] for l in layers)
        self._check(self.lib.b200_sn_power_iter_multi(_ptr(table), len(layers), max_h, max_w, int(iters),
                                                      int(bool(do_iter)), C.c_float(eps), _stream()),
                    "b200_sn_power_iter_multi")

    def adam_multi(self, table, n_entries, step, lr, beta1, beta2, eps):
        """one Adam update of every chunk in the device table (b200_adam_entry[]); `step` is a 0-dim fp32 CUDA tensor"""
        self._check(self.lib.b200_adam_multi(_ptr(table), int(n_entries), _ptr(step), C.c_double(lr), C.c_double(beta1),
                                             C.c_double(beta2), C.c_double(eps), _stream()), "b200_adam_multi")

    # ---- fused step arithmetic (loss.cu) ------------------------------------------------------------------------
    def loss_bce_groups(self, x, n, groups, split_group, target, weight, scale, partials, counts, slot):
        grad = torch.empty_like(x)
        self._check(self.lib.b200_loss_bce_groups(_ptr(x), int(n), int(groups), int(split_group), _ptr(target), _ptr(weight),
                                                  C.c_float(scale), _ptr(grad), _ptr(partials), _ptr(counts), int(slot),
                                                  _stream()), "b200_loss_bce_groups")
        return grad

    def loss_ce_groups(self, x, label, n, groups, weight, scale, partials, counts, slot):
        grad = torch.empty_like(x)
        self._check(self.lib.b200_loss_ce_groups(_ptr(x), _ptr(label), int(n), int(groups), int(x.shape[1]), _ptr(weight),
                                                 C.c_float(scale), _ptr(grad), _ptr(partials), _ptr(counts), int(slot),
                                                 _stream()), "b200_loss_ce_groups")
        return grad

    def loss_bce_pw_rows(self, x, t, sel, n, groups, n_sel, pos_weight, weight, scale, partials, counts, slot):
        grad = torch.empty_like(x)
        self._check(self.lib.b200_loss_bce_pw_rows(_ptr(x), _ptr(t), _ptr(sel), int(n), int(groups), int(x.shape[1]), int(n_sel),
                                                   _ptr(pos_weight), _ptr(weight), C.c_float(scale), _ptr(grad), _ptr(partials),
                                                   _ptr(counts), int(slot), _stream()), "b200_loss_bce_pw_rows")
        return grad

    def loss_l1_rows(self, a, b, N, L, b_stride_n, mask, denom, scale, partials, counts, slot):
        grad = torch.empty_like(a)
        self._check(self.lib.b200_loss_l1_rows(_ptr(a), _ptr(b), int(N), C.c_int64(L), C.c_int64(b_stride_n), _ptr(mask),
                                               C.c_float(denom), C.c_float(scale), _ptr(grad), _ptr(partials), _ptr(counts),
                                               int(slot), _stream()), "b200_loss_l1_rows")
        return grad

    def loss_kl(self, mu, logvar, scale, partials, counts, slot):
        dmu, dlv = torch.empty_like(mu), torch.empty_like(logvar)
        self._check(self.lib.b200_loss_kl(_ptr(mu), _ptr(logvar), C.c_int64(mu.numel()), C.c_float(scale), _ptr(dmu), _ptr(dlv),
                                          _ptr(partials), _ptr(counts), int(slot), _stream()), "b200_loss_kl")
        return dmu, dlv

    def loss_total(self, partials, counts, n_terms):
        terms = torch.empty((n_terms + 1,), dtype=torch.float32, device=partials.device)
        self._check(self.lib.b200_loss_total(_ptr(partials), _ptr(counts), int(n_terms), _ptr(terms), _stream()),
                    "b200_loss_total")
        return terms

    def copy_into(self, dst, dst_row, src):
        """dst[dst_row : dst_row + src.shape[0]] = src (contiguous tensors of equal row size and dtype)"""
        if not (dst.is_cuda and src.is_cuda):
            raise B200Error("copy_into: CUDA tensors required")
        if dst.dtype != src.dtype or dst[0].numel() != src[0].numel() or not (dst.is_contiguous() and src.is_contiguous()):
            raise B200Error("copy_into: mismatched row layout")
        rowb = dst[0].numel() * dst.element_size()
        self._check(self.lib.b200_copy(C.c_void_p(dst.data_ptr() + int(dst_row) * rowb), _ptr(src),
                                       C.c_size_t(src.shape[0] * rowb), _stream()), "b200_copy")


_K: Optional[Kernels] = None


class _Lazy:
    """Resolves to the loaded library on first attribute access (import of the package stays cheap and CPU-safe)."""

    def __getattr__(self, name):
        global _K
        if _K is None:
            _K = Kernels()
        return getattr(_K, name)


K = _Lazy()
