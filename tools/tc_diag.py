"""GPU diagnostic: tcgen05 conv kernels (fwd / dgrad / wgrad) vs an exact CPU reference on bf16-rounded operands."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "attribute-guided-image-generation-from-layout_b200"))
import torch
import torch.nn.functional as F
from b200gan import ops

def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))

def bf(t): return t.bfloat16().float()

CASES = [(64, 128, 4, 2, 1, 32, 3), (64, 128, 4, 2, 1, 33, 2), (128, 64, 3, 1, 1, 8, 5), (64, 64, 1, 1, 0, 16, 2),
         (64, 3, 7, 1, 3, 16, 2), (192, 256, 3, 1, 1, 8, 2), (128, 128, 5, 1, 2, 16, 1), (256, 512, 4, 2, 1, 8, 4),
         (64, 64, 3, 1, 1, 64, 4), (512, 1024, 4, 2, 1, 4, 8)]
ops.set_precision("bf16")
for (Cx, Cy, k, s, p, H, N) in CASES:
    g = torch.Generator().manual_seed(Cx + Cy + k)
    x = torch.randn(N, Cx, H, H, generator=g)
    w = torch.randn(Cy, Cx, k, k, generator=g) / (Cx * k * k) ** 0.5
    geom = ops.ConvGeom(Cx, Cy, k, k, s, p)
    packs = ops.WeightPacks()
    xd = x.permute(0, 2, 3, 1).contiguous().cuda()
    wd = w.cuda()
    y = ops.conv_forward(geom, packs, wd, xd, "cl", "cl")
    yr = F.conv2d(bf(x), bf(w), None, stride=s, padding=p)
    e_f = rel(y.permute(0, 3, 1, 2), yr)
    gy = torch.randn(yr.shape, generator=g)
    gyd = gy.permute(0, 2, 3, 1).contiguous().cuda()
    dx = ops.conv_dgrad(geom, packs, wd, gyd, "cl", (H, H), "cl")
    dxr = torch.nn.grad.conv2d_input(x.shape, bf(w), bf(gy), stride=s, padding=p)
    e_d = rel(dx.permute(0, 3, 1, 2), dxr)
    dw = torch.empty_like(wd)
    ops.conv_wgrad(geom, xd, "cl", gyd, "cl", dw)
    dwr = torch.nn.grad.conv2d_weight(bf(x), w.shape, bf(gy), stride=s, padding=p)
    e_w = rel(dw, dwr)
    torch.cuda.synchronize()
    print("Cx=%4d Cy=%4d k=%d s=%d H=%3d N=%d | fwd %.2e  dgrad %.2e  wgrad %.2e" % (Cx, Cy, k, s, H, N, e_f, e_d, e_w), flush=True)
