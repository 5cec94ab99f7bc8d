"""Box -> layout helpers on the device: the rasterised box masks and the shifted boxes
that the reference's data loader produces on the CPU (data/vg_custom_mask.py:120,136-158).  Bit-exact with the loader's
Python arithmetic; lets a caller ship 16 bytes per box to the GPU instead of two (O,1,H,W) fp32 mask tensors."""
import torch

from . import _lib


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
