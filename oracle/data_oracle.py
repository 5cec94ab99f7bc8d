"""TEST INFRASTRUCTURE — CPU restatement of the reference's data contract either side of the training step:
`imagenet_deprocess_batch` (data/utils.py:32-66, pinned by tests/golden/data.pt which the unmodified reference function
produced), the loader's one-hot attributes (data/vg_custom_mask.py:160-171) and `vg_collate_fn` (:176-221).  Box
rasterisation and shifted boxes live in gan_oracle.py (rasterize_boxes / shift_boxes).  Only tests/ import this."""
import torch

from . import gan_oracle as O

IMAGENET_MEAN = [0.485, 0.456, 0.406]
IMAGENET_STD = [0.229, 0.224, 0.225]


def imagenet_deprocess_batch(imgs: torch.Tensor, rescale: bool = True) -> torch.Tensor:
    """data/utils.py:47-66: Normalize(0, 1/std) -> Normalize(-mean, 1) -> per-image (x - min)/(max - min) -> *255 -> clamp
    -> byte, per image, fp32"""
    inv_std = torch.tensor([1.0 / s for s in IMAGENET_STD], dtype=torch.float32).view(3, 1, 1)
    neg_mean = torch.tensor([-m for m in IMAGENET_MEAN], dtype=torch.float32).view(3, 1, 1)
    out = []
    for img in imgs.detach().cpu().clone():
        x = img.sub(torch.zeros(3, 1, 1)).div(inv_std)
        x = x.sub(neg_mean).div(torch.ones(3, 1, 1))
        if rescale:
            lo, hi = x.min(), x.max()
            x = x.sub(lo).div(hi - lo)
        out.append(x.mul(255).clamp(0, 255).byte()[None])
    return torch.cat(out, dim=0)


def one_hot_attributes(att_idx: torch.Tensor, n_attributes: int) -> torch.Tensor:
    """vg_custom_mask.py:160-171: the row's indices up to (not including) the first -1 are switched on"""
    out = torch.zeros(att_idx.shape[0], n_attributes)
    for i in range(att_idx.shape[0]):
        n = 0
        while n < att_idx.shape[1] and att_idx[i][n] != -1:
            n += 1
        if n > 0:
            out[i, :] = torch.zeros(1, n_attributes).scatter_(1, att_idx[i, :].narrow(0, 0, n).unsqueeze(0), 1)
    return out


def collate(samples, n_attributes: int):
    """__getitem__'s mask / shift / one-hot work (vg_custom_mask.py:117-173) + vg_collate_fn (:176-221) for samples
    (image, objs, boxes, att_idx)"""
    H, W = samples[0][0].shape[-2:]
    imgs = torch.cat([s[0][None] for s in samples])
    objs = torch.cat([s[1] for s in samples])
    boxes = torch.cat([s[2] for s in samples])
    obj_to_img = torch.cat([torch.LongTensor(s[1].size(0)).fill_(i) for i, s in enumerate(samples)])
    attribute = torch.cat([one_hot_attributes(s[3], n_attributes) for s in samples])
    boxes_shift = O.shift_boxes(boxes)
    return (imgs, objs, boxes, O.rasterize_boxes(boxes, H, W), obj_to_img, attribute, O.rasterize_boxes(boxes_shift, H, W),
            boxes_shift)
