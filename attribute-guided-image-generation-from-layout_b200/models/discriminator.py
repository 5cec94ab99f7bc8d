"""Discriminators — module surface of the reference's models/discriminator.py on libb200gan kernels.

add_sn, OptimizedBlock, ResidualBlock, ImageDiscriminator, ObjectDiscriminator, AttributeDiscriminator and
AttributeDiscriminator128 keep the reference's names, constructor signatures, forward signatures / return values and
state_dict keys (discriminator.py:15-278).  Inputs are NCHW images / crops; logits come back as plain fp32 tensors.
"""
import torch
import torch.nn as nn

from b200gan import ops
from b200gan import nn as bnn
from b200gan.nn import add_sn  # noqa: F401  (discriminator.py:15-22)


class OptimizedBlock(nn.Module):
    """discriminator.py:29-60: pool?(conv3(ReLU(conv3(x)))) + sc1x1(pool?(x)); x is the NCHW network input"""

    def __init__(self, dim_in, dim_out, downsample=False):
        super().__init__()
        self.downsample = downsample
        self.resi = nn.Sequential(
            bnn.Conv2d(dim_in, dim_out, kernel_size=3, stride=1, padding=1, bias=True),
            bnn.ReLU(inplace=True),
            bnn.Conv2d(dim_out, dim_out, kernel_size=3, stride=1, padding=1, bias=True))
        self.learnable_sc = (dim_in != dim_out) or downsample
        if self.learnable_sc:
            self.sc = bnn.Conv2d(dim_in, dim_out, kernel_size=1, padding=0, bias=True)

    def forward(self, x, groups=1, out_relu=False):
        """out_relu: return relu(block output) — every consumer of a discriminator block applies ReLU first (the next
        ResidualBlock's in-place ReLU, or the trunk's final one), so the trunk asks for the activated tensor directly"""
        h = self.resi[0](x, x_layout="nchw", relu=True, groups=groups, grad_premasked=True)
        # the only consumer of the ReLU output above; with downsampling the pooling is folded into the convolution
        pooled_conv = self.downsample and ops.POOLED_CONV
        h = self.resi[2](h, groups=groups, mask_input_grad=True, pool=pooled_conv)
        s = x
        if self.downsample:
            if not pooled_conv:
                h = ops.avg_pool2(h)
            s = ops.pool_nchw(x, 2, 0.25)
        s = self.sc(s, x_layout="nchw", groups=groups)
        return ops.add_relu(h, s) if out_relu else ops.add(h, s)


class ResidualBlock(nn.Module):
    """discriminator.py:63-99.  resi[0] is an in-place ReLU evaluated before the shortcut, so BOTH branches consume
    relu(x).  Average pooling is linear, so pool(conv3(h)) + pool(sc(r)) is evaluated as
    conv4x4/2(h; fold(W)) + sc(pool(r)): the second convolution runs as one stride-2 convolution with the folded weight
    (16/36 of the multiply-adds, no full-resolution output) and the 1x1 shortcut on the pooled input (1/4 of its work).
    ops.POOLED_CONV = False restores the literal order (convolutions at full resolution, one pooling pass over h + s)."""

    def __init__(self, dim_in, dim_out, downsample=False):
        super().__init__()
        self.downsample = downsample
        self.resi = nn.Sequential(
            bnn.ReLU(inplace=True),
            bnn.Conv2d(dim_in, dim_in, kernel_size=3, stride=1, padding=1, bias=True),
            bnn.ReLU(inplace=True),
            bnn.Conv2d(dim_in, dim_out, kernel_size=3, stride=1, padding=1, bias=True))
        self.learnable_sc = (dim_in != dim_out) or downsample
        if self.learnable_sc:
            self.sc = bnn.Conv2d(dim_in, dim_out, kernel_size=1, padding=0, bias=True)

    def forward(self, x, groups=1, in_relu=False, out_relu=False):
        """in_relu: x already is relu(previous block output); out_relu: return relu(block output) (see OptimizedBlock)"""
        r = x if in_relu else ops.relu(x)
        if self.downsample and ops.POOLED_CONV:
            # the shortcut (pool, 1x1 convolution) is independent of the residual branch; ops.FORK_BRANCHES puts it on a
            # forked stream (off by default: measured 0.5 ms slower per step) — otherwise `forked` is a no-op
            with ops.forked(r) as fk:
                s = ops.avg_pool2(r)
                if self.learnable_sc:
                    s = self.sc(s, groups=groups)
            h = self.resi[1](r, relu=True, groups=groups, grad_premasked=True)
            h = self.resi[3](h, groups=groups, mask_input_grad=True, pool=True)
            fk.join()
            return ops.add_relu(h, s) if out_relu else ops.add(h, s)
        h = self.resi[1](r, relu=True, groups=groups, grad_premasked=True)
        h = self.resi[3](h, groups=groups, mask_input_grad=True)      # the only consumer of the ReLU output above
        s = self.sc(r, groups=groups) if self.learnable_sc else r
        if self.downsample:
            return ops.avg_pool2_sum(h, s, relu=out_relu)
        return ops.add_relu(h, s) if out_relu else ops.add(h, s)


def _trunk(net, x, groups):
    """`groups` calls of the network batched along dim 0: every spectral-normalised layer runs `groups` power
    iterations first (in call order) and call g's rows are scaled by its own 1/sigma_g."""
    bnn.sn_prepare(net, groups)
    h = x
    for i, blk in enumerate(net.main):
        # block outputs are consumed only through ReLU (the next block's in-place ReLU, the final one below):
        # each block hands over the activated tensor, fused into its last kernel
        h = blk(h, groups=groups, out_relu=True) if i == 0 else blk(h, groups=groups, in_relu=True, out_relu=True)
    N, H, W, C = h.shape
    return ops.pool(h, H, 1.0).view(N, C)     # in-place ReLU then sum over (H, W) (discriminator.py:224-226)


class AttributeDiscriminator128(nn.Module):
    """discriminator.py:102-141"""

    def __init__(self, conv_dim=64, downsample_first=False, n_attribute=128):
        super().__init__()
        self.main = nn.Sequential(
            OptimizedBlock(3, conv_dim, downsample=downsample_first),
            ResidualBlock(conv_dim, conv_dim * 2, downsample=True),
            ResidualBlock(conv_dim * 2, conv_dim * 4, downsample=True),
            ResidualBlock(conv_dim * 4, conv_dim * 8, downsample=True),
            ResidualBlock(conv_dim * 8, conv_dim * 16, downsample=True),
            ResidualBlock(conv_dim * 16, conv_dim * 16, downsample=True))
        self.classifier_att = bnn.Linear(conv_dim * 16, n_attribute)

    def forward(self, x, groups=1):
        return self.classifier_att(_trunk(self, x, groups), groups=groups, out_dtype=torch.float32)


class AttributeDiscriminator(nn.Module):
    """discriminator.py:144-181"""

    def __init__(self, conv_dim=64, downsample_first=False, n_attribute=128):
        super().__init__()
        self.main = nn.Sequential(
            OptimizedBlock(3, conv_dim, downsample=downsample_first),
            ResidualBlock(conv_dim, conv_dim * 2, downsample=True),
            ResidualBlock(conv_dim * 2, conv_dim * 4, downsample=True),
            ResidualBlock(conv_dim * 4, conv_dim * 8, downsample=True),
            ResidualBlock(conv_dim * 8, conv_dim * 16, downsample=True))
        self.classifier_att = bnn.Linear(conv_dim * 16, n_attribute)

    def forward(self, x, groups=1):
        return self.classifier_att(_trunk(self, x, groups), groups=groups, out_dtype=torch.float32)


class ImageDiscriminator(nn.Module):
    """discriminator.py:184-230"""

    def __init__(self, conv_dim=64):
        super().__init__()
        self.ch = conv_dim
        self.main = nn.Sequential(
            OptimizedBlock(3, self.ch, downsample=True),
            ResidualBlock(self.ch, self.ch * 2, downsample=True),
            ResidualBlock(self.ch * 2, self.ch * 4, downsample=True),
            ResidualBlock(self.ch * 4, self.ch * 8, downsample=True),
            ResidualBlock(self.ch * 8, self.ch * 16, downsample=True))
        self.classifier = bnn.Linear(self.ch * 16, 1, bias=False)

    def forward(self, x, groups=1):
        return self.classifier(_trunk(self, x, groups), groups=groups, out_dtype=torch.float32).view(-1)


class ObjectDiscriminator(nn.Module):
    """discriminator.py:233-278"""

    def __init__(self, conv_dim=64, n_class=0, downsample_first=False, n_attribute=128):
        super().__init__()
        self.main = nn.Sequential(
            OptimizedBlock(3, conv_dim, downsample=downsample_first),
            ResidualBlock(conv_dim, conv_dim * 2, downsample=True),
            ResidualBlock(conv_dim * 2, conv_dim * 4, downsample=True),
            ResidualBlock(conv_dim * 4, conv_dim * 8, downsample=True),
            ResidualBlock(conv_dim * 8, conv_dim * 16, downsample=True))
        self.classifier_src = bnn.Linear(conv_dim * 16, 1)
        self.classifier_cls = bnn.Linear(conv_dim * 16, n_class)

    def forward(self, x, y=None, groups=1):
        h = _trunk(self, x, groups)
        return (self.classifier_src(h, groups=groups, out_dtype=torch.float32).view(-1),
                self.classifier_cls(h, groups=groups, out_dtype=torch.float32))
