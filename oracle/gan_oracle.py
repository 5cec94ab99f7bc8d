"""CPU oracle for the G+D hot path (TEST INFRASTRUCTURE — never imported by the product).

A plain-PyTorch fp32, CPU, functional restatement of the reference's algorithm for one
generator + discriminator training step of the attribute-guided layout-to-image GAN.
It is written against a flat ``state`` dict whose keys and shapes equal the reference
modules' ``state_dict()``, so the same weights can be loaded into the
reference classes (in the build container, where /root/reference exists) and into the
B200 modules.  Every function cites the reference file:line it restates.

Pinning: the reference ships no tests or golden vectors.  The oracle is
pinned against outputs of the reference itself run in the build container; the generating
script is ``tests/golden/make_golden.py`` and the vectors live in ``tests/golden/*.pt``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / --impl
reference legs may import this module.
"""
from __future__ import annotations

import math
import zlib
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

State = Dict[str, torch.Tensor]

BN_EPS = 1e-5
BN_MOMENTUM = 0.1
SN_EPS = 1e-12

NUM_OBJECTS = 179      # data/vocab.json: object_idx_to_name
NUM_ATTRIBUTES = 106   # data/vocab.json: attribute_idx_to_name


# ------------------------------------------------------------------------------------
# state construction (shapes restate the reference constructors; values: PyTorch default
# init distributions drawn from a per-key generator so the dict is order independent)
# ------------------------------------------------------------------------------------

def _gen(seed: int, key: str) -> torch.Generator:
    g = torch.Generator()
    g.manual_seed((seed * 1000003 + zlib.crc32(key.encode())) & 0x7FFFFFFF)
    return g


def _uniform(shape, bound, g):
    return (torch.rand(shape, generator=g) * 2.0 - 1.0) * bound


class _Builder:
    def __init__(self, seed: int):
        self.seed = seed
        self.sd: State = {}

    def conv(self, name, cout, cin, k, bias, transpose=False):
        # nn.Conv2d / nn.ConvTranspose2d default init: kaiming_uniform(a=sqrt(5)) == U(+-1/sqrt(fan_in))
        shape = (cin, cout, k, k) if transpose else (cout, cin, k, k)
        fan_in = shape[1] * k * k
        b = 1.0 / math.sqrt(fan_in)
        self.sd[name + ".weight"] = _uniform(shape, b, _gen(self.seed, name + ".weight"))
        if bias:
            self.sd[name + ".bias"] = _uniform((cout,), b, _gen(self.seed, name + ".bias"))

    def linear(self, name, cout, cin, bias=True):
        b = 1.0 / math.sqrt(cin)
        self.sd[name + ".weight"] = _uniform((cout, cin), b, _gen(self.seed, name + ".weight"))
        if bias:
            self.sd[name + ".bias"] = _uniform((cout,), b, _gen(self.seed, name + ".bias"))

    def bn(self, name, c, affine=True):
        if affine:
            self.sd[name + ".weight"] = torch.ones(c)
            self.sd[name + ".bias"] = torch.zeros(c)
        self.sd[name + ".running_mean"] = torch.zeros(c)
        self.sd[name + ".running_var"] = torch.ones(c)
        self.sd[name + ".num_batches_tracked"] = torch.zeros((), dtype=torch.long)

    def cbn(self, name, c, classes):
        # generator_obj_att.py:31-38 — BN(affine=False) + Embedding(classes, 2C): gamma~N(1,.02), beta=0
        self.bn(name + ".bn", c, affine=False)
        w = torch.zeros(classes, 2 * c)
        w[:, :c] = 1.0 + 0.02 * torch.randn(classes, c, generator=_gen(self.seed, name + ".embed.weight"))
        self.sd[name + ".embed.weight"] = w

    def embedding(self, name, n, d):
        self.sd[name + ".weight"] = torch.randn(n, d, generator=_gen(self.seed, name + ".weight"))

    def spade(self, name, norm_nc, label_nc):
        # models/spade/networks/normalization.py:66-92
        self.bn(name + ".param_free_norm", norm_nc, affine=False)
        self.conv(name + ".mlp_shared.0", 128, label_nc, 3, True)
        self.conv(name + ".mlp_gamma", norm_nc, 128, 3, True)
        self.conv(name + ".mlp_beta", norm_nc, 128, 3, True)


def make_generator_state(seed: int = 0, image_size: int = 64, z_dim: int = 64, obj_att_dim: int = 64,
                         num_embeddings: int = NUM_OBJECTS, attribute_dim: int = NUM_ATTRIBUTES) -> State:
    """Shapes of models/generator_obj_att.py:603-616 (64) / generator_obj_att128.py:635-648 (128)."""
    b = _Builder(seed)
    cd = 64
    # CropEncoder generator_obj_att.py:367-393
    b.conv("crop_encoder.c1", cd, 3, 7, False)
    b.cbn("crop_encoder.bn1", cd, num_embeddings)
    for i, (name, bn) in enumerate((("c2", "bn2"), ("c3", "bn3"), ("c4", "bn4"), ("conv5", "bn5"))):
        cin = cd * 2 ** i
        b.conv("crop_encoder." + name, cin * 2, cin, 4, False)
        b.cbn("crop_encoder." + bn, cin * 2, num_embeddings)
    b.linear("crop_encoder.fc_mu", z_dim, cd * 16)
    b.linear("crop_encoder.fc_logvar", z_dim, cd * 16)
    # LayoutEncoder generator_obj_att.py:449-484
    hid = [128, 64, 64]
    cin = 512
    for i, h in enumerate(hid):
        b.conv("layout_encoder.clstm.cell_list.%d.conv" % i, 4 * h, cin + h, 5, True)
        cin = h
    for i in range(6):
        p = "layout_encoder.residual.%d.main." % i
        b.conv(p + "0", 64, 64, 3, False)
        b.bn(p + "1", 64)
        b.conv(p + "3", 64, 64, 3, False)
        b.bn(p + "4", 64)
    b.conv("layout_encoder.c0", cd, obj_att_dim + z_dim, 1, False)
    b.cbn("layout_encoder.bn1", cd, num_embeddings)
    for i, (name, bn) in enumerate((("c2", "bn2"), ("c3", "bn3"), ("c4", "bn4"))):
        c = cd * 2 ** i
        b.conv("layout_encoder." + name, c * 2, c, 4, False)
        b.cbn("layout_encoder." + bn, c * 2, num_embeddings)
    # Decoder generator_obj_att.py:516-544 (+ generator_obj_att128.py:549-557)
    b.conv("decoder.c0_new", cd * 4, cd + 128, 3, False)
    b.spade("decoder.spade_0", cd * 4, 64)
    b.conv("decoder.dc1", cd * 4, cd * 4, 4, False, transpose=True)
    b.spade("decoder.spade_1", cd * 4, 64)
    b.conv("decoder.dc2", cd * 2, cd * 4, 4, False, transpose=True)
    b.spade("decoder.spade_2", cd * 2, 64)
    b.conv("decoder.dc3", cd, cd * 2, 4, False, transpose=True)
    b.spade("decoder.spade_3", cd, 64)
    b.conv("decoder.c4", 3, cd, 7, True)
    if image_size == 128:
        b.conv("decoder.c5", cd * 2, 3, 7, False)
        b.spade("decoder.spade_4", cd * 2, 64)
        b.conv("decoder.c6", cd * 2, cd * 2, 5, False)
        b.spade("decoder.spade_5", cd * 2, 64)
        b.conv("decoder.c7", 3, cd * 2, 7, True)
    # GlobalEncoder generator_obj_att.py:425-435
    b.conv("global_encoder.c1", 128, 64, 4, False)
    b.bn("global_encoder.bn1", 128)
    b.conv("global_encoder.c2", 128, 128, 4, False)
    # AttributeEncoder generator_obj_att.py:575-586
    b.embedding("attribute_encoder.embedding", num_embeddings, obj_att_dim)
    b.linear("attribute_encoder.c0", 128, attribute_dim + obj_att_dim)
    b.bn("attribute_encoder.bn0", 128)
    b.linear("attribute_encoder.c1", 64, 128)
    b.bn("attribute_encoder.bn1", 64)
    b.linear("attribute_encoder.c2", 64, 64)
    return b.sd


def _sn_wrap(sd: State, name: str, seed: int):
    """torch.nn.utils.spectral_norm at wrap time (discriminator.py:15-22): weight -> weight_orig,
    u ~ normalize(N(0,1)) of size Cout, v ~ normalize(N(0,1)) of size Cin*kh*kw."""
    w = sd.pop(name + ".weight")
    sd[name + ".weight_orig"] = w
    h, wd = w.shape[0], w[0].numel()
    u = torch.randn(h, generator=_gen(seed, name + ".weight_u"))
    v = torch.randn(wd, generator=_gen(seed, name + ".weight_v"))
    sd[name + ".weight_u"] = F.normalize(u, dim=0, eps=SN_EPS)
    sd[name + ".weight_v"] = F.normalize(v, dim=0, eps=SN_EPS)


def make_discriminator_state(kind: str, seed: int = 0, conv_dim: int = 64, n_class: int = NUM_OBJECTS,
                             n_attribute: int = NUM_ATTRIBUTES, sn: bool = True) -> State:
    """kind in {'image','object','att','att128'}; shapes of discriminator.py:184-220, 233-262, 144-168, 102-128."""
    b = _Builder(seed + {"image": 11, "object": 22, "att": 33, "att128": 44}[kind])
    chans = [conv_dim, conv_dim * 2, conv_dim * 4, conv_dim * 8, conv_dim * 16]
    if kind == "att128":
        chans.append(conv_dim * 16)
    names = []
    b.conv("main.0.resi.0", chans[0], 3, 3, True)
    b.conv("main.0.resi.2", chans[0], chans[0], 3, True)
    b.conv("main.0.sc", chans[0], 3, 1, True)
    names += ["main.0.resi.0", "main.0.resi.2", "main.0.sc"]
    for i in range(1, len(chans)):
        cin, cout = chans[i - 1], chans[i]
        p = "main.%d." % i
        b.conv(p + "resi.1", cin, cin, 3, True)
        b.conv(p + "resi.3", cout, cin, 3, True)
        b.conv(p + "sc", cout, cin, 1, True)
        names += [p + "resi.1", p + "resi.3", p + "sc"]
    if kind == "image":
        b.linear("classifier", 1, chans[-1], bias=False)
        names.append("classifier")
    elif kind == "object":
        b.linear("classifier_src", 1, chans[-1])
        b.linear("classifier_cls", n_class, chans[-1])
        names += ["classifier_src", "classifier_cls"]
    else:
        b.linear("classifier_att", n_attribute, chans[-1])
        names.append("classifier_att")
    sd = b.sd
    if sn:
        for n in names:
            _sn_wrap(sd, n, b.seed)
    # state_dict ordering of the reference: bias, weight_orig, weight_u, weight_v per layer; order is
    # irrelevant for load_state_dict, so no reordering is attempted here.
    return sd


def clone_state(sd: State, requires_grad: bool = False) -> State:
    out = {}
    for k, v in sd.items():
        t = v.detach().clone()
        if requires_grad and t.is_floating_point() and is_parameter(k):
            t.requires_grad_(True)
        out[k] = t
    return out


def is_parameter(key: str) -> bool:
    return not (key.endswith("running_mean") or key.endswith("running_var") or key.endswith("num_batches_tracked")
                or key.endswith("weight_u") or key.endswith("weight_v"))


# ------------------------------------------------------------------------------------
# box crops (models/bilinear.py)
# ------------------------------------------------------------------------------------

def tensor_linspace(start: torch.Tensor, end: torch.Tensor, steps: int) -> torch.Tensor:
    """bilinear.py:252-281 — out[..., j] = sw[j]*start + ew[j]*end, sw=linspace(1,0), ew=linspace(0,1)
    (both built on CPU in fp32, bilinear.py:272-275); three separate fp32 roundings."""
    sw = torch.linspace(1, 0, steps=steps).to(start)
    ew = torch.linspace(0, 1, steps=steps).to(start)
    return sw * start.unsqueeze(-1) + ew * end.unsqueeze(-1)


def crop_grid(bbox: torch.Tensor, HH: int, WW: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """bilinear.py:125-131 — normalised sample coordinates X (B,WW), Y (B,HH) in [-1,1]."""
    bb = 2 * bbox - 1
    X = tensor_linspace(bb[:, 0], bb[:, 2], WW)
    Y = tensor_linspace(bb[:, 1], bb[:, 3], HH)
    return X, Y


def crop_taps(bbox: torch.Tensor, H: int, W: int, HH: int, WW: int):
    """Integer part of the crop (bit-exact contract): for align_corners=False,
    ATen GridSampler.h grid_sampler_unnormalize: ix = ((x + 1) * W - 1) / 2; taps floor(ix), floor(ix)+1.
    Returns ix0 (B,WW) int32, iy0 (B,HH) int32 and the fp32 fractional weights wx1=ix-ix0, wy1=iy-iy0."""
    X, Y = crop_grid(bbox.float(), HH, WW)
    ix = ((X + 1) * W - 1) / 2
    iy = ((Y + 1) * H - 1) / 2
    ix0 = torch.floor(ix)
    iy0 = torch.floor(iy)
    return ix0.to(torch.int32), iy0.to(torch.int32), ix - ix0, iy - iy0


def crop_bbox_batch(feats: torch.Tensor, bbox: torch.Tensor, bbox_to_feats: torch.Tensor, HH: int,
                    WW: Optional[int] = None) -> torch.Tensor:
    """bilinear.py:26-41,67-104,107-136 — crops[b] = grid_sample(feats[bbox_to_feats[b]], grid(bbox[b])),
    bilinear, zero padding, align_corners=False (torch>=1.3 default)."""
    if WW is None:
        WW = HH
    B = bbox.size(0)
    X, Y = crop_grid(bbox, HH, WW)
    grid = torch.stack([X.view(B, 1, WW).expand(B, HH, WW), Y.view(B, HH, 1).expand(B, HH, WW)], dim=3)
    src = feats[bbox_to_feats.to(feats.device)]
    return F.grid_sample(src, grid, mode="bilinear", padding_mode="zeros", align_corners=False)


def rasterize_boxes(boxes: torch.Tensor, H: int, W: int) -> torch.Tensor:
    """data/vg_custom_mask.py:120,136,157 — masks[i,:,round(y0*H):round(y1*H), round(x0*W):round(x1*W)] = 1
    with Python round (half-to-even) on double."""
    O = boxes.size(0)
    masks = torch.zeros(O, 1, H, W)
    for i in range(O):
        x0, y0, x1, y1 = [float(v) for v in boxes[i].tolist()]
        masks[i, :, round(y0 * H):round(y1 * H), round(x0 * W):round(x1 * W)] = 1
    return masks


def shift_boxes(boxes: torch.Tensor) -> torch.Tensor:
    """data/vg_custom_mask.py:139-158 — move narrow boxes 0.8x towards the farther horizontal border."""
    out = boxes.clone()
    for i in range(boxes.size(0)):
        x0, y0, x1, y1 = [float(v) for v in boxes[i].tolist()]
        if x1 - x0 < 0.5:
            left, right = x0, 1 - x1
            if left > right:
                s = left * 0.8
                x0, x1 = x0 - s, x1 - s
            elif right > left:
                s = right * 0.8
                x0, x1 = x0 + s, x1 + s
        out[i] = torch.tensor([x0, y0, x1, y1])
    return out


# ------------------------------------------------------------------------------------
# normalisation helpers
# ------------------------------------------------------------------------------------

def _bn(sd: State, prefix: str, x: torch.Tensor, training: bool, affine: bool) -> torch.Tensor:
    """nn.BatchNorm{1,2}d forward: batch stats (biased var) in training, running stats in eval;
    running_var uses the unbiased estimate; momentum 0.1; eps 1e-5."""
    w = sd[prefix + ".weight"] if affine else None
    b = sd[prefix + ".bias"] if affine else None
    if training:
        sd[prefix + ".num_batches_tracked"] += 1
    return F.batch_norm(x, sd[prefix + ".running_mean"], sd[prefix + ".running_var"], w, b, training,
                        BN_MOMENTUM, BN_EPS)


def _cbn(sd: State, prefix: str, x: torch.Tensor, y: torch.Tensor, training: bool) -> torch.Tensor:
    """ConditionalBatchNorm2d.forward generator_obj_att.py:40-44."""
    C = x.size(1)
    out = _bn(sd, prefix + ".bn", x, training, affine=False)
    emb = F.embedding(y, sd[prefix + ".embed.weight"])
    gamma, beta = emb[:, :C], emb[:, C:]
    return gamma.reshape(-1, C, 1, 1) * out + beta.reshape(-1, C, 1, 1)


def _spade(sd: State, prefix: str, x: torch.Tensor, seg: torch.Tensor, training: bool) -> torch.Tensor:
    """SPADE.forward models/spade/networks/normalization.py:94-108."""
    normalized = _bn(sd, prefix + ".param_free_norm", x, training, affine=False)
    seg = F.interpolate(seg, size=x.shape[2:], mode="nearest")
    actv = F.relu(F.conv2d(seg, sd[prefix + ".mlp_shared.0.weight"], sd[prefix + ".mlp_shared.0.bias"], padding=1))
    gamma = F.conv2d(actv, sd[prefix + ".mlp_gamma.weight"], sd[prefix + ".mlp_gamma.bias"], padding=1)
    beta = F.conv2d(actv, sd[prefix + ".mlp_beta.weight"], sd[prefix + ".mlp_beta.bias"], padding=1)
    return normalized * (1 + gamma) + beta


# ------------------------------------------------------------------------------------
# generator blocks
# ------------------------------------------------------------------------------------

def crop_encoder(sd: State, crops: torch.Tensor, objs: torch.Tensor, training: bool, z_dim: int,
                 eps: Optional[torch.Tensor] = None):
    """CropEncoder.forward generator_obj_att.py:395-422.  eps=None draws torch.randn from the CPU global
    RNG exactly like get_z_random (generator_obj_att.py:10-15,419)."""
    p = "crop_encoder."
    x = F.conv2d(crops, sd[p + "c1.weight"], None, stride=1, padding=3)
    x = F.relu(_cbn(sd, p + "bn1", x, objs, training))
    for conv, bn in (("c2", "bn2"), ("c3", "bn3"), ("c4", "bn4"), ("conv5", "bn5")):
        x = F.conv2d(x, sd[p + conv + ".weight"], None, stride=2, padding=1)
        x = F.relu(_cbn(sd, p + bn, x, objs, training))
    x = x.mean(dim=(2, 3))
    mu = F.linear(x, sd[p + "fc_mu.weight"], sd[p + "fc_mu.bias"])
    logvar = F.linear(x, sd[p + "fc_logvar.weight"], sd[p + "fc_logvar.bias"])
    std = (logvar * 0.5).exp()
    if eps is None:
        eps = torch.randn(std.size(0), std.size(1))
    z = eps.to(std) * std + mu
    return z, mu, logvar


def attribute_encoder(sd: State, objs: torch.Tensor, attribute: torch.Tensor, training: bool) -> torch.Tensor:
    """AttributeEncoder.forward generator_obj_att.py:588-600."""
    p = "attribute_encoder."
    a = torch.cat((F.embedding(objs, sd[p + "embedding.weight"]), attribute), dim=1)
    a = F.relu(_bn(sd, p + "bn0", F.linear(a, sd[p + "c0.weight"], sd[p + "c0.bias"]), training, True))
    a = F.relu(_bn(sd, p + "bn1", F.linear(a, sd[p + "c1.weight"], sd[p + "c1.bias"]), training, True))
    return F.linear(a, sd[p + "c2.weight"], sd[p + "c2.bias"])


def segment_lengths(obj_to_img: torch.Tensor) -> List[int]:
    """Run lengths of obj_to_img as LayoutConvLSTM.forward splits them (generator_obj_att.py:286-304):
    a new sequence starts whenever the image id changes."""
    ids = obj_to_img.tolist()
    lens, prev = [], ids[0] if ids else 0
    # the reference starts with previous_img_id = 0 (generator_obj_att.py:288); an initial id != 0 would
    # stack an empty list and raise, so a first id of 0 is part of the contract.
    cnt = 0
    for v in ids:
        if v == prev:
            cnt += 1
        else:
            lens.append(cnt)
            cnt, prev = 1, v
    lens.append(cnt)
    return lens


def conv_lstm(sd: State, prefix: str, x: torch.Tensor, obj_to_img: torch.Tensor, hidden=(128, 64, 64)) -> torch.Tensor:
    """LayoutConvLSTM.forward + ConvLSTMCell.forward generator_obj_att.py:271-346, 99-114: per image, the
    sequence of its objects runs through 3 stacked cells (gate order i,f,o,g; zero initial state); output
    is the last layer's final h.  Restated with the input-to-gate convolution hoisted over the sequence."""
    outs = []
    start = 0
    for n in segment_lengths(obj_to_img):
        seq = x[start:start + n]
        start += n
        for li, hid in enumerate(hidden):
            w = sd["%s.cell_list.%d.conv.weight" % (prefix, li)]
            b = sd["%s.cell_list.%d.conv.bias" % (prefix, li)]
            cin = seq.size(1)
            pre_x = F.conv2d(seq, w[:, :cin], b, padding=2)
            h = torch.zeros(1, hid, seq.size(2), seq.size(3), dtype=x.dtype, device=x.device)
            c = torch.zeros_like(h)
            hs = []
            for t in range(n):
                cc = pre_x[t:t + 1] + F.conv2d(h, w[:, cin:], None, padding=2)
                i, f, o, g = torch.split(cc, hid, dim=1)
                c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
                h = torch.sigmoid(o) * torch.tanh(c)
                hs.append(h)
            seq = torch.cat(hs, dim=0)
        outs.append(seq[-1:])
    return torch.cat(outs, dim=0)


def layout_encoder(sd: State, objs_att, masks, obj_to_img, z, objs, training: bool, image_size: int):
    """LayoutEncoder.forward generator_obj_att.py:487-513 (128: generator_obj_att128.py:489-511)."""
    p = "layout_encoder."
    e = torch.cat((objs_att, z), dim=1)
    h = e.view(e.size(0), e.size(1), 1, 1) * masks
    h = F.conv2d(h, sd[p + "c0.weight"], None, stride=1, padding=1)
    h = F.relu(_cbn(sd, p + "bn1", h, objs, training))
    h = F.relu(_cbn(sd, p + "bn2", F.conv2d(h, sd[p + "c2.weight"], None, stride=2, padding=1), objs, training))
    h = F.relu(_cbn(sd, p + "bn3", F.conv2d(h, sd[p + "c3.weight"], None, stride=2, padding=1), objs, training))
    h = _cbn(sd, p + "bn4", F.conv2d(h, sd[p + "c4.weight"], None, stride=2, padding=1), objs, training)
    if image_size == 128:
        h = F.adaptive_avg_pool2d(h, 8)
    h = conv_lstm(sd, p + "clstm", h, obj_to_img)
    for i in range(6):
        q = p + "residual.%d.main." % i
        r = F.conv2d(h, sd[q + "0.weight"], None, padding=1)
        r = F.relu(_bn(sd, q + "1", r, training, True))
        r = F.conv2d(r, sd[q + "3.weight"], None, padding=1)
        r = _bn(sd, q + "4", r, training, True)
        h = h + r
    return h


def global_encoder(sd: State, h: torch.Tensor, training: bool) -> torch.Tensor:
    """GlobalEncoder.forward generator_obj_att.py:437-446."""
    p = "global_encoder."
    h = F.conv2d(h, sd[p + "c1.weight"], None, stride=2, padding=1)
    h = F.relu(_bn(sd, p + "bn1", h, training, True))
    h = F.conv2d(h, sd[p + "c2.weight"], None, stride=2, padding=1)
    return h.sum(dim=(2, 3))


def decoder(sd: State, hidden: torch.Tensor, global_h: torch.Tensor, training: bool, image_size: int):
    """Decoder.forward generator_obj_att.py:546-572 (128: generator_obj_att128.py:560-604). No tanh."""
    p = "decoder."
    seg = hidden
    g = global_h[:, :, None, None].expand(-1, -1, 8, 8)
    h = F.conv2d(torch.cat((hidden, g), dim=1), sd[p + "c0_new.weight"], None, padding=1)
    h = F.relu(_spade(sd, p + "spade_0", h, seg, training))
    for i in (1, 2, 3):
        h = F.conv_transpose2d(h, sd[p + "dc%d.weight" % i], None, stride=2, padding=1)
        h = F.relu(_spade(sd, p + "spade_%d" % i, h, seg, training))
    h = F.conv2d(h, sd[p + "c4.weight"], sd[p + "c4.bias"], padding=3)
    if image_size == 128:
        up = F.interpolate(h, scale_factor=2, mode="nearest")
        h = F.conv2d(up, sd[p + "c5.weight"], None, padding=3)
        h = F.relu(_spade(sd, p + "spade_4", h, seg, training))
        h = F.conv2d(h, sd[p + "c6.weight"], None, padding=2)
        h = F.relu(_spade(sd, p + "spade_5", h, seg, training))
        h = F.conv2d(h, sd[p + "c7.weight"], sd[p + "c7.bias"], padding=3)
    return h


def generator_forward(sd: State, imgs, objs, boxes, masks, obj_to_img, z_rand, attribute, masks_shift, boxes_shift,
                      attribute_est, training: bool = True, image_size: int = 64, obj_size: Optional[int] = None,
                      eps: Optional[List[torch.Tensor]] = None):
    """Generator.forward generator_obj_att.py:618-647.  eps (optional) = the three CropEncoder noise draws
    in call order (input, rand, shift); None reproduces the reference's CPU global-RNG draws."""
    if obj_size is None:
        obj_size = image_size // 2
    z_dim = sd["crop_encoder.fc_mu.weight"].size(0)
    e = eps if eps is not None else [None, None, None]
    crops_input = crop_bbox_batch(imgs, boxes, obj_to_img, obj_size)
    z_rec, mu, logvar = crop_encoder(sd, crops_input, objs, training, z_dim, e[0])
    objs_att = attribute_encoder(sd, objs, attribute, training)
    objs_att_est = attribute_encoder(sd, objs, attribute_est, training)
    h_rec = layout_encoder(sd, objs_att_est, masks, obj_to_img, z_rec, objs, training, image_size)
    h_rand = layout_encoder(sd, objs_att, masks, obj_to_img, z_rand, objs, training, image_size)
    h_shift = layout_encoder(sd, objs_att, masks_shift, obj_to_img, z_rand, objs, training, image_size)
    g_rec = global_encoder(sd, h_rec, training)
    g_rand = global_encoder(sd, h_rand, training)
    g_shift = global_encoder(sd, h_shift, training)
    img_rec = decoder(sd, h_rec, g_rec, training, image_size)
    img_rand = decoder(sd, h_rand, g_rand, training, image_size)
    img_shift = decoder(sd, h_shift, g_shift, training, image_size)
    crops_rand = crop_bbox_batch(img_rand, boxes, obj_to_img, obj_size)
    _, z_rand_rec, _ = crop_encoder(sd, crops_rand, objs, training, z_dim, e[1])
    crops_input_rec = crop_bbox_batch(img_rec, boxes, obj_to_img, obj_size)
    crops_shift = crop_bbox_batch(img_shift, boxes_shift, obj_to_img, obj_size)
    _, z_rand_shift, _ = crop_encoder(sd, crops_shift, objs, training, z_dim, e[2])
    return (crops_input, crops_input_rec, crops_rand, crops_shift, img_rec, img_rand, img_shift, mu, logvar,
            z_rand_rec, z_rand_shift)


# ------------------------------------------------------------------------------------
# discriminators
# ------------------------------------------------------------------------------------

def sn_weight(sd: State, name: str, training: bool) -> torch.Tensor:
    """torch.nn.utils.spectral_norm.compute_weight (hooked by add_sn, discriminator.py:15-22): in train mode
    one power iteration, in place on u and v under no_grad; sigma = u . (W v); weight = weight_orig / sigma."""
    w = sd[name + ".weight_orig"]
    u, v = sd[name + ".weight_u"], sd[name + ".weight_v"]
    wm = w.reshape(w.size(0), -1)
    if training:
        with torch.no_grad():
            v.copy_(F.normalize(torch.mv(wm.t(), u), dim=0, eps=SN_EPS))
            u.copy_(F.normalize(torch.mv(wm, v), dim=0, eps=SN_EPS))
    sigma = torch.dot(u.detach().clone(), torch.mv(wm, v.detach().clone()))
    return w / sigma


def _dconv(sd: State, name: str, x: torch.Tensor, training: bool, sn: bool, padding: int) -> torch.Tensor:
    w = sn_weight(sd, name, training) if sn else sd[name + ".weight"]
    return F.conv2d(x, w, sd[name + ".bias"], padding=padding)


def _dlinear(sd: State, name: str, x: torch.Tensor, training: bool, sn: bool) -> torch.Tensor:
    w = sn_weight(sd, name, training) if sn else sd[name + ".weight"]
    return F.linear(x, w, sd.get(name + ".bias"))


def discriminator_trunk(sd: State, x: torch.Tensor, training: bool, sn: bool, downsample_first: bool) -> torch.Tensor:
    """main = OptimizedBlock + ResidualBlocks (discriminator.py:29-99), final in-place ReLU and sum over H,W.
    Hook order follows module call order: resi convs first, then sc.  ResidualBlock's shortcut consumes
    relu(x) because resi[0] is an in-place ReLU evaluated first."""
    h = _dconv(sd, "main.0.resi.0", x, training, sn, 1)
    h = _dconv(sd, "main.0.resi.2", F.relu(h), training, sn, 1)
    s = x
    if downsample_first:
        h = F.avg_pool2d(h, 2)
        s = F.avg_pool2d(x, 2)
    h = h + _dconv(sd, "main.0.sc", s, training, sn, 0)
    i = 1
    while ("main.%d.sc.bias" % i) in sd:
        p = "main.%d." % i
        r = F.relu(h)
        a = _dconv(sd, p + "resi.1", r, training, sn, 1)
        a = _dconv(sd, p + "resi.3", F.relu(a), training, sn, 1)
        a = F.avg_pool2d(a, 2)
        s = F.avg_pool2d(_dconv(sd, p + "sc", r, training, sn, 0), 2)
        h = a + s
        i += 1
    return F.relu(h).sum(dim=(2, 3))


def image_discriminator(sd: State, x, training=True, sn=True):
    """ImageDiscriminator.forward discriminator.py:222-230."""
    h = discriminator_trunk(sd, x, training, sn, downsample_first=True)
    return _dlinear(sd, "classifier", h, training, sn).view(-1)


def object_discriminator(sd: State, x, training=True, sn=True):
    """ObjectDiscriminator.forward discriminator.py:264-278."""
    h = discriminator_trunk(sd, x, training, sn, downsample_first=False)
    src = _dlinear(sd, "classifier_src", h, training, sn)
    cls = _dlinear(sd, "classifier_cls", h, training, sn)
    return src.view(-1), cls


def attribute_discriminator(sd: State, x, training=True, sn=True):
    """AttributeDiscriminator[128].forward discriminator.py:170-181, 130-141."""
    h = discriminator_trunk(sd, x, training, sn, downsample_first=False)
    return _dlinear(sd, "classifier_att", h, training, sn)


# ------------------------------------------------------------------------------------
# synthetic VG-shaped batch and the train step (train64.py:141-370)
# ------------------------------------------------------------------------------------

def synth_batch(n_images: int, image_size: int = 64, objs_per_image: Optional[int] = 8, seed: int = 0,
                sparse_attributes: bool = False) -> Dict[str, torch.Tensor]:
    g = torch.Generator()
    g.manual_seed(seed)
    imgs = torch.randn(n_images, 3, image_size, image_size, generator=g)
    counts = [objs_per_image if objs_per_image else int(torch.randint(3, 10, (1,), generator=g)) for _ in range(n_images)]
    O = sum(counts)
    obj_to_img = torch.cat([torch.full((c,), i, dtype=torch.long) for i, c in enumerate(counts)])
    objs = torch.randint(1, NUM_OBJECTS, (O,), generator=g)
    xy0 = torch.rand(O, 2, generator=g) * 0.6
    wh = torch.rand(O, 2, generator=g) * 0.3 + 0.1
    xy1 = torch.clamp(xy0 + wh, max=1.0)
    boxes = torch.cat([xy0, xy1], dim=1).float()
    boxes_shift = shift_boxes(boxes)
    masks = rasterize_boxes(boxes, image_size, image_size)
    masks_shift = rasterize_boxes(boxes_shift, image_size, image_size)
    attribute = torch.zeros(O, NUM_ATTRIBUTES)
    n_att = torch.randint(1, 3, (O,), generator=g)
    for i in range(O):
        if sparse_attributes and i % 2 == 1:
            continue
        idx = torch.randperm(NUM_ATTRIBUTES, generator=g)[: int(n_att[i])]
        attribute[i, idx] = 1
    z = torch.randn(O, 64, generator=g)
    return dict(imgs=imgs, objs=objs, boxes=boxes, masks=masks, obj_to_img=obj_to_img, attribute=attribute,
                masks_shift=masks_shift, boxes_shift=boxes_shift, z=z)


def pos_weight_vector() -> torch.Tensor:
    """train64.py:25-28 — (100000 - count)/count; counts live in the reference's attribute_counts.py which
    cannot travel; the parity harness uses a fixed synthetic count table with the same formula."""
    counts = torch.arange(NUM_ATTRIBUTES, dtype=torch.float32) * 37.0 + 150.0
    return (100000.0 - counts) / counts


LAMBDAS = dict(img_adv=1.0, obj_adv=1.0, obj_cls=1.0, z_rec=8.0, img_rec=1.0, kl=0.01, att_cls=2.0)  # train64.py:439-446


def estimate_attributes(att_logits: torch.Tensor, attribute: torch.Tensor) -> torch.Tensor:
    """train64.py:155-166 as it behaves when every object is annotated or, for un-annotated rows, with the
    intended semantics (argmax attribute switched on).  (The literal code breaks on torch>=1.2)"""
    est = attribute.clone()
    none = attribute.sum(dim=1) == 0
    idx = att_logits.argmax(1)
    rows = none.nonzero().view(-1)
    est[rows, idx[rows]] = 1
    return est


def swap_attributes(attribute: torch.Tensor, attribute_est: torch.Tensor, objs: torch.Tensor, obj_to_img: torch.Tensor,
                    n_images: int, matrix: torch.Tensor, rng):
    """train64.py:169-188 "change GT attribute": in the first floor(N/3) images the first floor(n_obj/2) objects get 1-2
    NEW attributes drawn from the object-vs-attribute co-occurrence row `matrix[obj]` with the object's current attributes
    zeroed out; both `attribute` and `attribute_est` rows are replaced by the new multi-hot vector.  `rng` is a Python
    `random.Random` (the reference uses the global `random` module: randrange(1, 3) is drawn BEFORE choices(), as Python
    evaluates the keyword argument first).  Returns (attribute, attribute_est, swapped_rows); inputs are not modified;
    attribute_GT of the step stays the ORIGINAL attribute (train64.py:153)."""
    attribute, attribute_est = attribute.clone(), attribute_est.clone()
    attribute_GT = attribute.clone()
    rows = []
    for img_idx in range(n_images // 3):
        obj_indices = torch.nonzero(obj_to_img == img_idx).view(-1)
        for changed, obj_idx in enumerate(obj_indices.tolist()):
            if changed >= len(obj_indices) // 2:
                break
            old = torch.nonzero(attribute_GT[obj_idx]).view(-1)
            weights = matrix[int(objs[obj_idx])].clone()
            weights[old] = 0
            k = rng.randrange(1, 3)
            new = rng.choices(range(attribute.shape[1]), weights, k=k)
            attribute[obj_idx] = 0
            attribute[obj_idx, torch.tensor(new, dtype=torch.long)] = 1
            attribute_est[obj_idx] = attribute[obj_idx]
            rows.append(obj_idx)
    return attribute, attribute_est, torch.tensor(rows, dtype=torch.long)


def bce_logits(x, target_value: float):
    return F.binary_cross_entropy_with_logits(x, torch.full_like(x, target_value))


def d_step_loss(nets, batch, fake, pos_weight, lam=LAMBDAS):
    """train64.py:195-252.  nets = dict(image=fn, object=fn, att=fn) closures over their states;
    fake = generator outputs (detached here exactly as the reference does)."""
    (crops_input, crops_input_rec, crops_rand, crops_shift, img_rec, img_rand, img_shift) = [t.detach() for t in fake[:7]]
    d_img_fake = 0.4 * bce_logits(nets["image"](img_rec), 0) + 0.4 * bce_logits(nets["image"](img_rand), 0) \
        + 0.2 * bce_logits(nets["image"](img_shift), 0)
    d_img_real = bce_logits(nets["image"](batch["imgs"]), 1)
    d_obj_fake = 0.4 * bce_logits(nets["object"](crops_input_rec)[0], 0) + 0.4 * bce_logits(nets["object"](crops_rand)[0], 0) \
        + 0.2 * bce_logits(nets["object"](crops_shift)[0], 0)
    src, cls = nets["object"](crops_input)
    d_obj_real = bce_logits(src, 1)
    d_obj_cls = F.cross_entropy(cls, batch["objs"])
    att_cls = nets["att"](crops_input)
    att_idx = batch["attribute_GT"].sum(dim=1).nonzero().view(-1)
    d_att = F.binary_cross_entropy_with_logits(att_cls.index_select(0, att_idx), batch["attribute_GT"].index_select(0, att_idx),
                                               pos_weight=pos_weight)
    loss = lam["img_adv"] * (d_img_fake + d_img_real) + lam["obj_adv"] * (d_obj_fake + d_obj_real) \
        + lam["obj_cls"] * d_obj_cls + lam["att_cls"] * d_att
    return loss, dict(d_img_fake=d_img_fake, d_img_real=d_img_real, d_obj_fake=d_obj_fake, d_obj_real=d_obj_real,
                      d_obj_cls=d_obj_cls, d_att=d_att)


def g_step_loss(nets, batch, out, pos_weight, lam=LAMBDAS):
    """train64.py:284-364."""
    (crops_input, crops_input_rec, crops_rand, crops_shift, img_rec, img_rand, img_shift, mu, logvar, z_rand_rec,
     z_rand_shift) = out
    imgs, z, objs, attribute = batch["imgs"], batch["z"], batch["objs"], batch["attribute"]
    N = imgs.shape[0]
    n_change = N // 3
    rec_mask = torch.ones(N)
    rec_mask[:n_change] = 0
    g_img_rec = (rec_mask.to(imgs) * (img_rec - imgs).abs().view(N, -1).mean(1)).sum() / (N - n_change)
    g_z_rec = 0.5 * (z_rand_rec - z).abs().mean() + 0.5 * (z_rand_shift - z).abs().mean()
    g_kl = -0.5 * torch.sum(1 + logvar - mu.pow(2) - logvar.exp())
    g_img_adv = 0.4 * bce_logits(nets["image"](img_rec), 1) + 0.4 * bce_logits(nets["image"](img_rand), 1) \
        + 0.2 * bce_logits(nets["image"](img_shift), 1)
    att_idx = attribute.sum(dim=1).nonzero().view(-1)
    att_t = attribute.index_select(0, att_idx)
    adv, cls, att = [], [], []
    for crops in (crops_input_rec, crops_rand, crops_shift):
        src, c = nets["object"](crops)
        adv.append(bce_logits(src, 1))
        cls.append(F.cross_entropy(c, objs))
        a = nets["att"](crops)
        att.append(F.binary_cross_entropy_with_logits(a.index_select(0, att_idx), att_t, pos_weight=pos_weight))
    w = (0.4, 0.4, 0.2)
    g_obj_adv = sum(wi * v for wi, v in zip(w, adv))
    g_obj_cls = sum(wi * v for wi, v in zip(w, cls))
    g_obj_att = sum(wi * v for wi, v in zip(w, att))
    loss = lam["img_rec"] * g_img_rec + lam["z_rec"] * g_z_rec + lam["img_adv"] * g_img_adv + lam["obj_adv"] * g_obj_adv \
        + lam["obj_cls"] * g_obj_cls + lam["att_cls"] * g_obj_att + lam["kl"] * g_kl
    return loss, dict(g_img_rec=g_img_rec, g_z_rec=g_z_rec, g_kl=g_kl, g_img_adv=g_img_adv, g_obj_adv=g_obj_adv,
                      g_obj_cls=g_obj_cls, g_obj_att=g_obj_att)


class OracleModel:
    """The four nets of train64.py:99-109 as states + closures; step() restates train64.py:141-370 minus the
    optimizers' update (gradients are left in .grad), logging and checkpointing."""

    def __init__(self, image_size: int = 64, seed: int = 0, states: Optional[Dict[str, State]] = None):
        self.image_size = image_size
        self.obj_size = image_size // 2
        if states is None:
            states = make_states(image_size, seed)
        self.G = clone_state(states["G"], True)
        self.D_img = clone_state(states["D_img"], True)
        self.D_obj = clone_state(states["D_obj"], True)
        self.D_att = clone_state(states["D_att"], True)
        self.pos_weight = pos_weight_vector()
        self.training = True

    def nets(self):
        t = self.training
        return dict(image=lambda x: image_discriminator(self.D_img, x, t),
                    object=lambda x: object_discriminator(self.D_obj, x, t),
                    att=lambda x: attribute_discriminator(self.D_att, x, t))

    def generator(self, b, attribute_est, eps=None):
        return generator_forward(self.G, b["imgs"], b["objs"], b["boxes"], b["masks"], b["obj_to_img"], b["z"],
                                 b["attribute"], b["masks_shift"], b["boxes_shift"], attribute_est,
                                 training=self.training, image_size=self.image_size, obj_size=self.obj_size, eps=eps)

    def zero_grad(self, which):
        for st in which:
            for v in st.values():
                if v.requires_grad:
                    v.grad = None

    def step(self, batch, eps_d=None, eps_g=None):
        b = dict(batch)
        b["attribute_GT"] = b["attribute"].clone()
        nets = self.nets()
        with torch.no_grad():
            crops = crop_bbox_batch(b["imgs"], b["boxes"], b["obj_to_img"], self.obj_size)
        est_logits = nets["att"](crops)                                   # train64.py:160-161
        attribute_est = estimate_attributes(est_logits.detach(), b["attribute"])
        out_d = self.generator(b, attribute_est, eps_d)                      # train64.py:191
        d_loss, d_terms = d_step_loss(nets, b, out_d, self.pos_weight)
        self.zero_grad((self.D_img, self.D_obj, self.D_att))
        d_loss.backward()
        d_grads = {n: {k: v.grad.clone() for k, v in st.items() if v.requires_grad and v.grad is not None}
                   for n, st in (("D_img", self.D_img), ("D_obj", self.D_obj), ("D_att", self.D_att))}
        out_g = self.generator(b, attribute_est, eps_g)                      # train64.py:280
        g_loss, g_terms = g_step_loss(nets, b, out_g, self.pos_weight)
        self.zero_grad((self.G,))
        g_loss.backward()
        g_grads = {k: v.grad.clone() for k, v in self.G.items() if v.requires_grad and v.grad is not None}
        return dict(d_loss=d_loss.detach(), g_loss=g_loss.detach(), d_terms=d_terms, g_terms=g_terms, d_grads=d_grads,
                    g_grads=g_grads, out_d=[t.detach() for t in out_d], out_g=[t.detach() for t in out_g],
                    attribute_est=attribute_est)


def make_states(image_size: int = 64, seed: int = 0) -> Dict[str, State]:
    return dict(G=make_generator_state(seed, image_size),
                D_img=make_discriminator_state("image", seed),
                D_obj=make_discriminator_state("object", seed),
                D_att=make_discriminator_state("att128" if image_size == 128 else "att", seed))
