"""models/bilinear.py surface of the reference: differentiable box crops on libb200gan (b200_crop_fwd / b200_crop_bwd).

crop_bbox_batch(feats, bbox, bbox_to_feats, HH, WW=None, backend='cudnn') keeps the reference signature
(reference models/bilinear.py:26) and semantics: crops[b] = bilinear sample of feats[bbox_to_feats[b]] inside
bbox[b] on an inclusive-endpoint HH x WW grid, zero padding, align_corners=False (what F.grid_sample does under the
torch version the reference runs with here).  bbox_to_feats is the CPU LongTensor the training script
keeps on the host (train64.py:150); a CUDA tensor is accepted too (one D2H copy).
"""
import torch

from b200gan import ops


def crop_bbox_batch(feats, bbox, bbox_to_feats, HH, WW=None, backend='cudnn'):
    if backend != 'cudnn':
        raise NotImplementedError("only the default 'cudnn' (grid_sample) semantics of the reference are provided")
    assert bbox.size(0) == bbox_to_feats.size(0) and bbox.size(1) == 4
    return ops.crop_bbox_batch(feats, bbox, bbox_to_feats, HH, WW)


def crop_bbox_batch_cudnn(feats, bbox, bbox_to_feats, HH, WW=None):
    return crop_bbox_batch(feats, bbox, bbox_to_feats, HH, WW)


def crop_bbox(feats, bbox, HH, WW=None, backend='cudnn'):
    """reference models/bilinear.py:107 — one box per feature map"""
    idx = torch.arange(feats.size(0), dtype=torch.long)
    return crop_bbox_batch(feats, bbox, idx, HH, WW, backend)
