"""64x64 generator — module surface of the reference's models/generator_obj_att.py on libb200gan kernels.

Class names, constructor signatures, Generator.forward's signature / 11-tuple and every state_dict key follow the
reference (generator_obj_att.py:603-647).  Internally activations are channel-last fp32 and each op
is a torch.autograd.Function over the C ABI (b200gan/ops.py).  Differences in *how* (never in what):
  * LayoutEncoder's `embedding (x) mask -> 1x1 conv(pad 1)` is evaluated in its rank-1 form (W e) (x) mask
    (generator_obj_att.py:489-491), never materialising the (O,128,H,W) tensor;
  * LayoutConvLSTM batches all images per time step (time-major packing) and hoists the input-to-gate
    convolution out of the recurrence (generator_obj_att.py:271-346);
  * BN + (conditional / SPADE) modulation + ReLU / residual add are single fused kernels.
"""
import torch
import torch.nn as nn

from b200gan import ops
from b200gan import nn as bnn
from models.bilinear import crop_bbox_batch
from models.spade.networks.normalization import SPADE


def get_z_random(batch_size, z_dim, random_type='gauss'):
    """generator_obj_att.py:10-15 — drawn on the CPU global RNG (parity with the reference's RNG stream)"""
    if random_type == 'uni':
        return torch.rand(batch_size, z_dim) * 2.0 - 1.0
    return torch.randn(batch_size, z_dim)


class ConditionalBatchNorm2d(nn.Module):
    """generator_obj_att.py:31-44; forward(x, y): x (O,H,W,C) channel-last, y (O,) int32 class ids"""

    def __init__(self, num_features, num_classes):
        super().__init__()
        self.num_features = num_features
        self.bn = bnn.BatchNorm2d(num_features, affine=False)
        self.embed = bnn.Embedding(num_classes, num_features * 2)
        self.embed.weight.data[:, :num_features].normal_(1, 0.02)
        self.embed.weight.data[:, num_features:].zero_()

    def forward(self, x, y, relu=False, groups=1):
        if self.bn.training:
            bnn.count_batches(self.bn.num_batches_tracked, groups)
        return ops.cond_batch_norm(x, self.embed.weight, y, self.bn.running_mean, self.bn.running_var, self.bn.training,
                                   relu, groups)


class ResidualBlock(nn.Module):
    """generator_obj_att.py:47-60: x + BN(conv3(ReLU(BN(conv3(x)))))"""

    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.main = nn.Sequential(
            bnn.Conv2d(dim_in, dim_out, kernel_size=3, stride=1, padding=1, bias=False),
            bnn.BatchNorm2d(dim_out, affine=True, track_running_stats=True),
            bnn.ReLU(inplace=True),
            bnn.Conv2d(dim_out, dim_out, kernel_size=3, stride=1, padding=1, bias=False),
            bnn.BatchNorm2d(dim_out, affine=True, track_running_stats=True))

    def forward(self, x, groups=1):
        h = self.main[0](x, stats=True)
        h = self.main[1](h, relu=True, groups=groups)
        h = self.main[3](h, stats=True)
        return self.main[4](h, residual=x, groups=groups)


class ConvLSTMCell(nn.Module):
    """generator_obj_att.py:63-118 — holds the (4*hid, cin+hid, k, k) gate convolution; evaluated by LayoutConvLSTM"""

    def __init__(self, input_size, input_dim, hidden_dim, kernel_size, bias):
        super().__init__()
        self.height, self.width = input_size
        self.input_dim, self.hidden_dim = input_dim, hidden_dim
        self.kernel_size = kernel_size
        self.padding = kernel_size[0] // 2, kernel_size[1] // 2
        self.bias = bias
        self.conv = bnn.Conv2d(in_channels=input_dim + hidden_dim, out_channels=4 * hidden_dim,
                               kernel_size=kernel_size, padding=self.padding, bias=bias)
        self._layer = ops.ConvLSTMLayer(input_dim, hidden_dim, kernel_size[0])


class LayoutConvLSTM(nn.Module):
    """generator_obj_att.py:232-364; forward(obj_tensor (O,8,8,C) channel-last, obj_to_img CPU LongTensor)"""

    def __init__(self, input_size, input_dim, hidden_dim, kernel_size, bias=True, return_all_layers=False):
        super().__init__()
        if not isinstance(hidden_dim, (list, tuple)):
            hidden_dim = [hidden_dim]
        self.height, self.width = input_size, input_size
        self.input_dim, self.hidden_dim, self.num_layers = input_dim, list(hidden_dim), len(hidden_dim)
        cells = []
        for i, h in enumerate(self.hidden_dim):
            cin = input_dim if i == 0 else self.hidden_dim[i - 1]
            cells.append(ConvLSTMCell((input_size, input_size), cin, h, kernel_size, bias))
        self.cell_list = nn.ModuleList(cells)

    def forward(self, obj_tensor, obj_to_img, hidden_state=None):
        plan = ops.get_plan(obj_to_img, None, obj_tensor.device)
        params = []
        for c in self.cell_list:
            params += [c.conv.weight, c.conv.bias]
        return ops.conv_lstm(obj_tensor, plan, [c._layer for c in self.cell_list], params)


class CropEncoder(nn.Module):
    """generator_obj_att.py:367-422; forward(imgs (O,3,S,S) NCHW, objs int32) -> z, mu, logvar"""

    def __init__(self, conv_dim=64, z_dim=8, class_num=10):
        super().__init__()
        self.c1 = bnn.Conv2d(3, conv_dim, kernel_size=7, stride=1, padding=3, bias=False)
        self.bn1 = ConditionalBatchNorm2d(conv_dim, class_num)
        self.c2 = bnn.Conv2d(conv_dim, conv_dim * 2, kernel_size=4, stride=2, padding=1, bias=False)
        self.bn2 = ConditionalBatchNorm2d(conv_dim * 2, class_num)
        self.c3 = bnn.Conv2d(conv_dim * 2, conv_dim * 4, kernel_size=4, stride=2, padding=1, bias=False)
        self.bn3 = ConditionalBatchNorm2d(conv_dim * 4, class_num)
        self.c4 = bnn.Conv2d(conv_dim * 4, conv_dim * 8, kernel_size=4, stride=2, padding=1, bias=False)
        self.bn4 = ConditionalBatchNorm2d(conv_dim * 8, class_num)
        self.conv5 = bnn.Conv2d(conv_dim * 8, conv_dim * 16, kernel_size=4, stride=2, padding=1, bias=False)
        self.bn5 = ConditionalBatchNorm2d(conv_dim * 16, class_num)
        self.fc_mu = bnn.Linear(conv_dim * 16, z_dim)
        self.fc_logvar = bnn.Linear(conv_dim * 16, z_dim)
        self.eps_source = None   # optional callable(O, z_dim, device) -> device tensor (CUDA-graph friendly)

    def forward(self, imgs, objs, groups=1):
        """groups > 1: that many calls batched along dim 0 (imgs (groups*O,3,S,S), objs (groups*O,)); batch statistics,
        running-statistics updates and the noise draws stay per call, in call order."""
        x = self.bn1(self.c1(imgs, x_layout="nchw", stats=True), objs, relu=True, groups=groups)
        x = self.bn2(self.c2(x, stats=True), objs, relu=True, groups=groups)
        x = self.bn3(self.c3(x, stats=True), objs, relu=True, groups=groups)
        x = self.bn4(self.c4(x, stats=True), objs, relu=True, groups=groups)
        x = self.bn5(self.conv5(x, stats=True), objs, relu=True, groups=groups)
        O, H, W, C = x.shape
        x = ops.pool(x, H, 1.0 / (H * W)).view(O, C)
        mu = self.fc_mu(x, out_dtype=torch.float32)              # the VAE head and everything after it stay fp32
        logvar = self.fc_logvar(x, out_dtype=torch.float32)
        if self.eps_source is not None:
            eps = self.eps_source(O, mu.size(1), mu.device)
        elif groups == 1:
            eps = get_z_random(O, mu.size(1)).to(mu.device)
        else:
            eps = torch.cat([get_z_random(O // groups, mu.size(1)) for _ in range(groups)]).to(mu.device)
        z = ops.reparameterize(mu, logvar, eps)
        return z, mu, logvar


class GlobalEncoder(nn.Module):
    """generator_obj_att.py:425-446; (N,8,8,64) -> (N,128)"""

    def __init__(self):
        super().__init__()
        self.c1 = bnn.Conv2d(64, 128, kernel_size=4, stride=2, padding=1, bias=False)
        self.bn1 = bnn.BatchNorm2d(128)
        self.c2 = bnn.Conv2d(128, 128, kernel_size=4, stride=2, padding=1, bias=False)

    def forward(self, h, groups=1):
        h = self.bn1(self.c1(h, stats=True), relu=True, groups=groups)
        h = self.c2(h)
        N, H, W, C = h.shape
        return ops.pool(h, H, 1.0).view(N, C)


class LayoutEncoder(nn.Module):
    """generator_obj_att.py:449-513 (+ generator_obj_att128.py:486,505 pool when image_size == 128)"""

    def __init__(self, conv_dim=64, z_dim=8, obj_att_dim=64, class_num=10, resi_num=6, clstm_layers=3, att_dim=64,
                 pool_to_8=False):
        super().__init__()
        if clstm_layers == 1:
            self.clstm = LayoutConvLSTM(8, 512, [64], (5, 5))
        elif clstm_layers == 2:
            self.clstm = LayoutConvLSTM(8, 512, [64, 64], (5, 5))
        elif clstm_layers == 3:
            self.clstm = LayoutConvLSTM(8, 512, [128, 64, 64], (5, 5))
        self.residual = nn.Sequential(*[ResidualBlock(dim_in=64, dim_out=64) for _ in range(resi_num)])
        self.c0 = bnn.Conv2d(obj_att_dim + z_dim, conv_dim, kernel_size=1, stride=1, padding=1, bias=False)
        self.bn1 = ConditionalBatchNorm2d(conv_dim, class_num)
        self.c2 = bnn.Conv2d(conv_dim, conv_dim * 2, kernel_size=4, stride=2, padding=1, bias=False)
        self.bn2 = ConditionalBatchNorm2d(conv_dim * 2, class_num)
        self.c3 = bnn.Conv2d(conv_dim * 2, conv_dim * 4, kernel_size=4, stride=2, padding=1, bias=False)
        self.bn3 = ConditionalBatchNorm2d(conv_dim * 4, class_num)
        self.c4 = bnn.Conv2d(conv_dim * 4, conv_dim * 8, kernel_size=4, stride=2, padding=1, bias=False)
        self.bn4 = ConditionalBatchNorm2d(conv_dim * 8, class_num)
        self._pool_to_8 = pool_to_8
        self._c0_packs = ops.WeightPacks()

    def forward(self, objs_att, masks, obj_to_img, z, objs, groups=1):
        """groups > 1: that many calls batched along dim 0; obj_to_img must then number the images of call g from
        g * N on, so every call keeps its own ConvLSTM sequences."""
        e = ops.concat_channels(objs_att, z)                                       # (O, att + z)
        w0 = self.c0.weight
        v = ops.linear(e, w0.view(w0.shape[0], w0.shape[1]), None, self._c0_packs)   # W0 e  (rank-1 form of c0)
        h = ops.mask_outer(v, masks)                                               # (O,H+2,W+2,64), zero ring = padding 1
        h = self.bn1(h, objs, relu=True, groups=groups)
        h = self.bn2(self.c2(h, stats=True), objs, relu=True, groups=groups)
        h = self.bn3(self.c3(h, stats=True), objs, relu=True, groups=groups)
        h = self.bn4(self.c4(h, stats=True), objs, groups=groups)
        if self._pool_to_8:
            f = h.shape[1] // 8
            h = ops.pool(h, f, 1.0 / (f * f))
        h = self.clstm(h, obj_to_img)
        for blk in self.residual:
            h = blk(h, groups=groups)
        return h


class Decoder(nn.Module):
    """generator_obj_att.py:516-572 (+ generator_obj_att128.py:549-604 refinement when image_size == 128)"""

    def __init__(self, nf=64, conv_dim=64, image_size=64):
        super().__init__()
        self.sw, self.sh, self.h_dim = 8, 8, 64
        self.c0_new = bnn.Conv2d(conv_dim + 128, conv_dim * 4, kernel_size=3, stride=1, padding=1, bias=False)
        self.spade_0 = SPADE(conv_dim * 4, self.h_dim)
        self.dc1 = bnn.ConvTranspose2d(conv_dim * 4, conv_dim * 4, kernel_size=4, stride=2, padding=1, bias=False)
        self.spade_1 = SPADE(conv_dim * 4, self.h_dim)
        self.dc2 = bnn.ConvTranspose2d(conv_dim * 4, conv_dim * 2, kernel_size=4, stride=2, padding=1, bias=False)
        self.spade_2 = SPADE(conv_dim * 2, self.h_dim)
        self.dc3 = bnn.ConvTranspose2d(conv_dim * 2, conv_dim * 1, kernel_size=4, stride=2, padding=1, bias=False)
        self.spade_3 = SPADE(conv_dim * 1, self.h_dim)
        self.c4 = bnn.Conv2d(conv_dim * 1, 3, kernel_size=7, stride=1, padding=3, bias=True)
        self._refine = image_size == 128
        if self._refine:
            self.c5 = bnn.Conv2d(3, conv_dim * 2, kernel_size=7, stride=1, padding=3, bias=False)
            self.spade_4 = SPADE(conv_dim * 2, self.h_dim)
            self.c6 = bnn.Conv2d(conv_dim * 2, conv_dim * 2, kernel_size=5, stride=1, padding=2, bias=False)
            self.spade_5 = SPADE(conv_dim * 2, self.h_dim)
            self.c7 = bnn.Conv2d(conv_dim * 2, 3, kernel_size=7, stride=1, padding=3, bias=True)

    def forward(self, hidden, global_h, z=None, groups=1):
        """hidden (N,8,8,64) channel-last, global_h (N,128) -> image (N,3,S,S) NCHW (no tanh, as the reference)"""
        N, H, W, C = hidden.shape
        seg = hidden
        x = ops.concat_channels(hidden.view(N * H * W, C), global_h, 1, H * W).view(N, H, W, C + global_h.shape[1])
        h = self.c0_new(x, stats=True)
        h = self.spade_0.forward_cl(h, seg, relu=True, groups=groups)
        h = self.spade_1.forward_cl(self.dc1(h), seg, relu=True, groups=groups)
        h = self.spade_2.forward_cl(self.dc2(h), seg, relu=True, groups=groups)
        h = self.spade_3.forward_cl(self.dc3(h), seg, relu=True, groups=groups)
        img = self.c4(h, out_layout="nchw")
        if not self._refine:
            return img
        up = ops.upsample_nearest_nchw(img, 2)
        h = self.c5(up, x_layout="nchw", stats=True)
        h = self.spade_4.forward_cl(h, seg, relu=True, groups=groups)
        h = self.spade_5.forward_cl(self.c6(h, stats=True), seg, relu=True, groups=groups)
        return self.c7(h, out_layout="nchw")


class AttributeEncoder(nn.Module):
    """generator_obj_att.py:575-600"""

    def __init__(self, attribute_dim=106, embedding_dim=64, class_num=10):
        super().__init__()
        self.embedding = bnn.Embedding(class_num, embedding_dim)
        self.c0 = bnn.Linear(attribute_dim + embedding_dim, 128)
        self.bn0 = bnn.BatchNorm1d(128)
        self.c1 = bnn.Linear(128, 64)
        self.bn1 = bnn.BatchNorm1d(64)
        self.c2 = bnn.Linear(64, 64)

    def forward(self, objs, attribute, groups=1):
        a = ops.concat_channels(self.embedding(objs), attribute.contiguous())
        a = self.bn0(self.c0(a), relu=True, groups=groups)
        a = self.bn1(self.c1(a), relu=True, groups=groups)
        return self.c2(a)


class Generator(nn.Module):
    """generator_obj_att.py:603-647"""
    _image_size = 64

    def __init__(self, num_embeddings, obj_att_dim=64, z_dim=8, obj_size=64, clstm_layers=3, attribute_dim=128):
        super().__init__()
        self.obj_size = obj_size
        big = self._image_size == 128
        self.crop_encoder = CropEncoder(z_dim=z_dim, class_num=num_embeddings)
        self.layout_encoder = LayoutEncoder(z_dim=z_dim, obj_att_dim=obj_att_dim, class_num=num_embeddings,
                                            clstm_layers=clstm_layers, pool_to_8=big)
        self.decoder = Decoder(image_size=self._image_size)
        self.global_encoder = GlobalEncoder()
        self.attribute_encoder = AttributeEncoder(attribute_dim=attribute_dim, embedding_dim=obj_att_dim,
                                                  class_num=num_embeddings)

    def forward(self, imgs, objs, boxes, masks, obj_to_img, z_rand, attribute, masks_shift, boxes_shift, attribute_est):
        return self.forward_batched(imgs, objs, boxes, masks, obj_to_img, z_rand, attribute, masks_shift, boxes_shift,
                                    attribute_est)["outputs"]

    def forward_batched(self, imgs, objs, boxes, masks, obj_to_img, z_rand, attribute, masks_shift, boxes_shift,
                        attribute_est, defer_tail=False):
        """generator_obj_att.py:618-647 with the three passes (rec, rand, shift) of every sub-network batched along
        dim 0 as three groups: one launch serves all three, while batch statistics, running-statistics updates, noise
        draws and ConvLSTM sequences stay per pass and in the reference's call order.  Returns the reference's 11-tuple
        ("outputs") plus the batched tensors the training step feeds to the discriminators:
        "imgs_fake" (3N,3,H,W) = [img_rec; img_rand; img_shift], "crops_fake" (3O,3,S,S) = [crops_input_rec; crops_rand;
        crops_shift].

        defer_tail (training step only): the LAST sub-network call — crop_encoder on the generated crops, whose result only
        the z-reconstruction loss reads — runs on a forked stream, and the returned dict carries "join": a callable that
        makes the current stream wait for it.  The caller must call it before reading "mu2" / outputs[9:11], before the
        next call of this module and before the optimizers run; until then the call overlaps whatever the caller launches
        (the discriminator passes)."""
        o2i = obj_to_img.cpu() if obj_to_img.is_cuda else obj_to_img
        N, O = imgs.shape[0], objs.shape[0]
        objs32 = objs.to(torch.int32)
        objs32x3 = objs32.repeat(3)
        o2i3 = torch.cat([o2i, o2i + N, o2i + 2 * N])
        z_rand = z_rand.contiguous()
        crops_input = crop_bbox_batch(imgs, boxes, o2i, self.obj_size)
        z_rec, mu, logvar = self.crop_encoder(crops_input, objs32)

        # attribute_encoder(objs, attribute) then attribute_encoder(objs, attribute_est)  (generator_obj_att.py:622-623)
        att2 = self.attribute_encoder(objs32x3[:2 * O], torch.cat([attribute, attribute_est]), groups=2)
        objs_att, objs_att_est = att2[:O], att2[O:]

        # layout_encoder x3: (att_est, masks, z_rec), (att, masks, z_rand), (att, masks_shift, z_rand)   (:626-628)
        att3 = torch.cat([objs_att_est, objs_att, objs_att])
        z3 = torch.cat([z_rec, z_rand, z_rand])
        masks3 = torch.cat([masks, masks, masks_shift])
        h3 = self.layout_encoder(att3, masks3, o2i3, z3, objs32x3, groups=3)          # (3N,8,8,64)
        g3 = self.global_encoder(h3, groups=3)                                       # (3N,128)
        imgs_fake = self.decoder(h3, g3, groups=3)                                   # (3N,3,H,W): rec, rand, shift
        img_rec, img_rand, img_shift = imgs_fake[:N], imgs_fake[N:2 * N], imgs_fake[2 * N:]

        # crops of (img_rec, boxes), (img_rand, boxes), (img_shift, boxes_shift)   (:639-644)
        boxes3 = torch.cat([boxes, boxes, boxes_shift])
        crops_fake = crop_bbox_batch(imgs_fake, boxes3, o2i3, self.obj_size)
        crops_input_rec, crops_rand, crops_shift = crops_fake[:O], crops_fake[O:2 * O], crops_fake[2 * O:]
        # crop_encoder(crops_rand) then crop_encoder(crops_shift)   (:640, 645)
        join = None
        if defer_tail and crops_fake.is_cuda:
            cur = torch.cuda.current_stream(crops_fake.device)
            tail = ops.side_streams(crops_fake.device, 3)[2]
            tail.wait_stream(cur)
            with torch.cuda.stream(tail):
                _, mu2, _ = self.crop_encoder(crops_fake[O:], objs32x3[:2 * O], groups=2)

            def join():
                torch.cuda.current_stream(crops_fake.device).wait_stream(tail)
        else:
            _, mu2, _ = self.crop_encoder(crops_fake[O:], objs32x3[:2 * O], groups=2)
        z_rand_rec, z_rand_shift = mu2[:O], mu2[O:]
        outputs = (crops_input, crops_input_rec, crops_rand, crops_shift, img_rec, img_rand, img_shift, mu, logvar,
                   z_rand_rec, z_rand_shift)
        return dict(outputs=outputs, imgs_fake=imgs_fake, crops_fake=crops_fake, mu2=mu2, join=join)
