"""TEST INFRASTRUCTURE — CPU restatement of `masks_to_layout` (BASELINE.json config 5, north_star (2)).

**Parity unpinned to the reference tree**: the reference CALLS `masks_to_layout(vecs, boxes, masks, obj_to_img, H=..., N=...)`
(utils/draw_box.py:482-483) but its definition (`util.layout`, draw_box.py:22, commented import) is not shipped — SURVEY.md
F3.  The function comes from sg2im (google/sg2im, sg2im/layout.py, the code base the reference's data pipeline and
crop/layout utilities derive from, see data/vg_custom_mask.py header), whose published algorithm is restated here:

    grid   = boxes_to_grid(boxes, H, W)            X = (linspace(0,1,W) - x0) / (x1 - x0), Y likewise, grid = 2*[X,Y] - 1
    img_in = vecs.view(O,D,1,1) * masks.view(O,1,M,M)
    sampled = F.grid_sample(img_in, grid)          bilinear, zeros padding (torch 2.11 default align_corners=False)
    out[n] = sum over objects o of image n (scatter_add along dim 0) -> (N, D, H, W)

Only tests/, __graft_entry__.smoke() and tools/bench_layout.py may import this module."""
import torch
import torch.nn.functional as F


def boxes_to_grid(boxes: torch.Tensor, H: int, W: int) -> torch.Tensor:
    O = boxes.size(0)
    b = boxes.view(O, 4, 1, 1)
    x0, y0, x1, y1 = b[:, 0], b[:, 1], b[:, 2], b[:, 3]
    ww, hh = x1 - x0, y1 - y0
    X = torch.linspace(0, 1, steps=W).view(1, 1, W).to(boxes)
    Y = torch.linspace(0, 1, steps=H).view(1, H, 1).to(boxes)
    X = (X - x0) / ww
    Y = (Y - y0) / hh
    grid = torch.stack([X.expand(O, H, W), Y.expand(O, H, W)], dim=3)
    return grid.mul(2).sub(1)


def masks_to_layout(vecs, boxes, masks, obj_to_img, H, W=None, N=None):
    """vecs (O,D), boxes (O,4) [x0,y0,x1,y1] in [0,1], masks (O,M,M), obj_to_img (O,) -> (N,D,H,W), sum pooling"""
    O, D = vecs.shape
    M = masks.size(1)
    W = H if W is None else W
    N = int(obj_to_img.max()) + 1 if N is None else N
    grid = boxes_to_grid(boxes, H, W)
    img_in = vecs.view(O, D, 1, 1) * masks.float().view(O, 1, M, M)
    sampled = F.grid_sample(img_in, grid, mode="bilinear", padding_mode="zeros", align_corners=False)
    out = torch.zeros(N, D, H, W, dtype=vecs.dtype)
    idx = obj_to_img.view(O, 1, 1, 1).expand(O, D, H, W)
    return out.scatter_add(0, idx, sampled)


def layout_taps(boxes, M, H, W):
    """integer part of the sampling (bit-exact contract): per object and output column / row, the floor of the unnormalised
    source coordinate ((g + 1) * M - 1) / 2 and the fractional weight, fp32 in torch's operation order"""
    grid = boxes_to_grid(boxes, H, W)
    ix = ((grid[:, 0, :, 0] + 1) * M - 1) / 2          # (O, W)
    iy = ((grid[:, :, 0, 1] + 1) * M - 1) / 2          # (O, H)
    ix0, iy0 = torch.floor(ix), torch.floor(iy)
    return ix0.to(torch.int32), iy0.to(torch.int32), ix - ix0, iy - iy0
