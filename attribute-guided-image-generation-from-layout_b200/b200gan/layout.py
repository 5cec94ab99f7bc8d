"""Box -> layout helpers on the device: the rasterised box masks and the shifted boxes
that the reference's data loader produces on the CPU (data/vg_custom_mask.py:120,136-158).  Bit-exact with the loader's
Python arithmetic; lets a caller ship 16 bytes per box to the GPU instead of two (O,1,H,W) fp32 mask tensors."""
import torch

from . import _lib, ops


def rasterize_boxes(boxes: torch.Tensor, H: int, W: int) -> torch.Tensor:
    """boxes (O,4) fp32 [x0,y0,x1,y1] in [0,1] on the GPU -> masks (O,1,H,W) fp32 {0,1}"""
    return _lib.K.rasterize_boxes(boxes, int(H), int(W))


def shift_boxes(boxes: torch.Tensor) -> torch.Tensor:
    """vg_custom_mask.py:139-158: the boxes of the generator's "shift" pass"""
    return _lib.K.shift_boxes(boxes)


def layout_inputs(boxes: torch.Tensor, image_size: int):
    """(masks, boxes_shift, masks_shift) for a batch of boxes — what the loader hands to Generator.forward"""
    bs = shift_boxes(boxes)
    return rasterize_boxes(boxes, image_size, image_size), bs, rasterize_boxes(bs, image_size, image_size)


# ---- masks_to_layout (utils/draw_box.py:482-483; sg2im layout.py semantics, see oracle/layout_oracle.py) -------------------
_LIN = {}


def _linspace01(steps: int, device) -> torch.Tensor:
    """torch.linspace(0, 1, steps) built on the CPU in fp32 exactly as boxes_to_grid does, then uploaded (cached)"""
    key = (steps, str(device))
    t = _LIN.get(key)
    if t is None:
        t = _LIN[key] = torch.linspace(0, 1, steps=steps).to(device)
    return t


class _MasksToLayoutFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, vecs, boxes, masks, plan, H, W):
        vecs, boxes, masks = vecs.contiguous().float(), boxes.contiguous().float(), masks.contiguous().float()
        linx, liny = _linspace01(W, vecs.device), _linspace01(H, vecs.device)
        ctx.save_for_backward(vecs, boxes, masks, linx, liny)
        ctx.plan = plan
        return _lib.K.m2l_fwd(vecs, boxes, masks, plan.img_box_start, plan.box_order, linx, liny, plan.N)

    @staticmethod
    def backward(ctx, dout):
        vecs, boxes, masks, linx, liny = ctx.saved_tensors
        dvecs, dmasks = _lib.K.m2l_bwd(dout.contiguous(), vecs, boxes, masks, ctx.plan.box_to_img, linx, liny,
                                       ctx.needs_input_grad[0], ctx.needs_input_grad[2])
        return dvecs, None, dmasks, None, None, None


def masks_to_layout(vecs, boxes, masks, obj_to_img, H, W=None, N=None):
    """vecs (O,D), boxes (O,4) [x0,y0,x1,y1], masks (O,M,M), obj_to_img (O,) -> (N,D,H,W): every object's embedding, weighted
    by its mask bilinearly resampled into its box, summed per image (deterministic gather).  Differentiable w.r.t. vecs and
    masks.  obj_to_img may live on the CPU (as the reference keeps it) or the GPU; it need not be sorted."""
    W = H if W is None else W
    o2i = obj_to_img.cpu() if obj_to_img.is_cuda else obj_to_img
    n = (int(o2i.max()) + 1 if o2i.numel() else 0) if N is None else int(N)
    plan = ops.get_plan(o2i, n, vecs.device)
    return _MasksToLayoutFn.apply(vecs, boxes, masks, plan, int(H), int(W))


def masks_to_layout_taps(boxes, M, H, W=None):
    """(ix0 (O,W), iy0 (O,H), fx, fy): floors and fractions of the source coordinates — the bit-exact part of the contract"""
    W = H if W is None else W
    return _lib.K.m2l_taps(boxes.contiguous().float(), _linspace01(W, boxes.device), _linspace01(H, boxes.device), M)
