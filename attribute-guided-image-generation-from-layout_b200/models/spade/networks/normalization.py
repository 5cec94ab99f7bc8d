"""SPADE (reference models/spade/networks/normalization.py:66-108) on libb200gan.

out = BN(x, affine=False) * (1 + gamma(seg)) + beta(seg), seg nearest-resized to x, gamma/beta from a shared
3x3 conv + ReLU followed by two 3x3 convs.  Here the gamma and beta convolutions run as ONE gather-GEMM with 2C
outputs and the normalise/modulate(/ReLU) step is a single fused kernel (b200_norm_fwd, mode SPADE).
"""
import torch
import torch.nn as nn

from b200gan import ops
from b200gan import nn as bnn
from b200gan.ops import ConvGeom, WeightPacks


class SPADE(nn.Module):
    def __init__(self, norm_nc, label_nc):
        super().__init__()
        self.param_free_norm = bnn.BatchNorm2d(norm_nc, affine=False)
        nhidden = 128
        self.mlp_shared = nn.Sequential(bnn.Conv2d(label_nc, nhidden, kernel_size=3, padding=1), bnn.ReLU())
        self.mlp_gamma = bnn.Conv2d(nhidden, norm_nc, kernel_size=3, padding=1)
        self.mlp_beta = bnn.Conv2d(nhidden, norm_nc, kernel_size=3, padding=1)
        self._gb_geom = ConvGeom(nhidden, 2 * norm_nc, 3, 3, 1, 1)
        self._gb_packs = None

    def forward_cl(self, x, seg, relu=False, groups=1):
        """x (N,H,W,C), seg (N,h,w,label_nc) channel-last; H must be an integer multiple of h."""
        f = x.shape[1] // seg.shape[1]
        if f > 1:
            seg = ops.upsample_nearest(seg, f)
        actv = self.mlp_shared[0](seg, relu=True, grad_premasked=True)      # its only consumer masks the gradient (below)
        if self._gb_packs is None or self._gb_packs.sources[0] is not self.mlp_gamma.weight:
            self._gb_packs = WeightPacks(sources=(self.mlp_gamma.weight, self.mlp_beta.weight))
        w = torch.cat([self.mlp_gamma.weight, self.mlp_beta.weight], dim=0)
        b = torch.cat([self.mlp_gamma.bias, self.mlp_beta.bias], dim=0)
        gb = ops.conv2d(actv, w, b, self._gb_geom, self._gb_packs, mask_input_grad=True)
        bn = self.param_free_norm
        if bn.training:
            bnn.count_batches(bn.num_batches_tracked, groups)
        return ops.spade_norm(x, gb, bn.running_mean, bn.running_var, bn.training, relu, groups)

    def forward(self, x, segmap):
        """reference signature: NCHW in, NCHW out"""
        y = self.forward_cl(ops.nchw_to_cl(x), ops.nchw_to_cl(segmap))
        return ops.cl_to_nchw(y)
