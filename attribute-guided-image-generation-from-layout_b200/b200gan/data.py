"""Device side of the reference's data contract.

`collate_on_device` replaces `vg_collate_fn` + the mask / shifted-box / one-hot work of `VgSceneGraphDataset.__getitem__`
(data/vg_custom_mask.py:117-173, 176-221): the host ships 16 bytes per box, the attribute index lists and the images; the
(O,1,H,W) masks, the shifted boxes and masks and the (O,106) multi-hot attributes are produced on the GPU, bit-identical
to what the loader computes with Python arithmetic.  `imagenet_deprocess_batch` is data/utils.py:47-66 on the device.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch

from . import _lib

IMAGENET_MEAN = [0.485, 0.456, 0.406]          # data/utils.py:21-25
IMAGENET_STD = [0.229, 0.224, 0.225]
_CONST = {}


def _deprocess_constants(device):
    """fp32(1/std), fp32(-mean) exactly as data/utils.py:24-25 + torchvision's Normalize build them (Python double, then fp32)"""
    key = str(device)
    if key not in _CONST:
        _CONST[key] = (torch.tensor([1.0 / s for s in IMAGENET_STD], dtype=torch.float32, device=device),
                       torch.tensor([-m for m in IMAGENET_MEAN], dtype=torch.float32, device=device))
    return _CONST[key]


def imagenet_deprocess_batch(imgs: torch.Tensor, rescale: bool = True) -> torch.Tensor:
    """(N,3,H,W) fp32 normalised images on the GPU -> (N,3,H,W) uint8 in [0,255] (data/utils.py:47-66), on the device"""
    if imgs.dim() != 4 or imgs.shape[1] != 3:
        raise _lib.B200Error("imagenet_deprocess_batch: (N,3,H,W) images required")
    inv_std, neg_mean = _deprocess_constants(imgs.device)
    return _lib.K.imagenet_deprocess(imgs.detach().contiguous().float(), inv_std, neg_mean, rescale)


def one_hot_attributes(att_idx: torch.Tensor, n_attributes: int) -> torch.Tensor:
    """(O,A) int64 attribute index lists, -1 terminated (vg_custom_mask.py:160-171) -> (O,n_attributes) fp32 multi-hot"""
    return _lib.K.one_hot_attributes(att_idx.contiguous(), int(n_attributes))


Sample = Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]


def collate_on_device(samples: Sequence[Sample], n_attributes: int, device, pin: bool = True):
    """samples: per image (image (3,H,W) fp32, objs (O_i,) int64, boxes (O_i,4) fp32 [x0,y0,x1,y1], att_idx (O_i,A) int64 with
    -1 padding) on the CPU.  Returns the tuple `vg_collate_fn` yields — (imgs, objs, boxes, masks, obj_to_img, attribute,
    masks_shift, boxes_shift) — with everything on `device` except obj_to_img, which stays on the CPU like in the
    reference's loop (train64.py:150)."""
    from . import layout
    imgs = torch.stack([s[0] for s in samples])
    objs = torch.cat([s[1] for s in samples])
    boxes = torch.cat([s[2] for s in samples]).float()
    att = torch.cat([s[3] for s in samples])
    obj_to_img = torch.cat([torch.full((s[1].shape[0],), i, dtype=torch.long) for i, s in enumerate(samples)])
    H, W = imgs.shape[-2:]

    def up(t):
        return (t.pin_memory() if pin else t).to(device, non_blocking=True)

    imgs_d, objs_d, boxes_d, att_d = up(imgs), up(objs), up(boxes), up(att)
    masks, boxes_shift, masks_shift = layout.rasterize_boxes(boxes_d, H, W), None, None
    boxes_shift = layout.shift_boxes(boxes_d)
    masks_shift = layout.rasterize_boxes(boxes_shift, H, W)
    attribute = one_hot_attributes(att_d, n_attributes)
    return imgs_d, objs_d, boxes_d, masks, obj_to_img, attribute, masks_shift, boxes_shift
