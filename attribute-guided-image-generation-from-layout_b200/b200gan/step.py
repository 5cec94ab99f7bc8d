"""One G+D training step with the B200 modules — the wiring of the reference's train64.py:141-370 / train128.py.

`TrainStep` owns the four networks (Generator, Image/Object/Attribute discriminators with spectral norm) and their
Adam optimizers exactly as train64.py:99-114 builds them, and exposes `step(batch)` = attribute estimation, D-step
(forward, losses, backward, 3x Adam) and G-step (forward, losses, backward, Adam).  The loss arithmetic on the tiny
logit tensors stays in PyTorch; everything inside the networks runs libb200gan kernels.

Two things the reference computes and then throws away are skipped (and subtracted from the FLOP numerator in
bench.py): the autograd graph of the D-step's generator forward (all its uses are detached, train64.py:195-240) and
the discriminators' weight gradients during the G-step (zeroed before use, train64.py:254-256).
"""
from __future__ import annotations

import math
import os
from typing import Dict, Optional

import torch
import torch.nn.functional as F

from . import ops

DEFER_TAIL = os.environ.get("B200_DEFER_TAIL", "1") != "0"       # the generator's last crop_encoder call next to the D passes
OVERLAP_G2 = os.environ.get("B200_OVERLAP_G2", "1") != "0"       # the G-step's generator pass next to the D-step
PARALLEL_D = os.environ.get("B200_PARALLEL_D", "1") != "0"      # the three discriminators on forked streams (TrainStep._side_by_side)

LAMBDAS = dict(img_adv=1.0, obj_adv=1.0, obj_cls=1.0, z_rec=8.0, img_rec=1.0, kl=0.01, att_cls=2.0)  # train64.py:439-446
NUM_OBJECTS, NUM_ATTRIBUTES = 179, 106                                                               # data/vocab.json


def default_pos_weight() -> torch.Tensor:
    """(100000 - count) / count (train64.py:25-28) over a synthetic count table; the real table is
    attribute_counts.py in the reference tree and is passed in by callers that have it."""
    counts = torch.arange(NUM_ATTRIBUTES, dtype=torch.float32) * 37.0 + 150.0
    return (100000.0 - counts) / counts


def build_networks(image_size: int = 64, z_dim: int = 64, embedding_dim: int = 64):
    """train64.py:99-109 / train128.py:100-110"""
    if image_size == 128:
        from models.generator_obj_att128 import Generator
        from models.discriminator import AttributeDiscriminator128 as AttributeDiscriminator
    else:
        from models.generator_obj_att import Generator
        from models.discriminator import AttributeDiscriminator
    from models.discriminator import ImageDiscriminator, ObjectDiscriminator, add_sn
    netG = Generator(num_embeddings=NUM_OBJECTS, obj_att_dim=embedding_dim, z_dim=z_dim, clstm_layers=3,
                     obj_size=image_size // 2, attribute_dim=NUM_ATTRIBUTES)
    netD_image = add_sn(ImageDiscriminator(conv_dim=embedding_dim))
    netD_object = add_sn(ObjectDiscriminator(n_class=NUM_OBJECTS))
    netD_att = add_sn(AttributeDiscriminator(n_attribute=NUM_ATTRIBUTES))
    return netG, netD_image, netD_object, netD_att


def bce_const(logits, value: float):
    return F.binary_cross_entropy_with_logits(logits, torch.full_like(logits, value))


def bce_const_groups(logits, value: float, groups: int):
    """per-call means of a batched call: (groups,)"""
    l = F.binary_cross_entropy_with_logits(logits, torch.full_like(logits, value), reduction="none")
    return l.view(groups, -1).mean(1)


FAKE_W = (0.4, 0.4, 0.2)          # rec, rand, shift weights (train64.py:199-201, 219-221, 298-349)


def estimate_attributes(att_logits, attribute):
    """train64.py:155-166 (intended semantics): un-annotated objects get their arg-max attribute switched on."""
    none = (attribute.sum(dim=1, keepdim=True) == 0).to(attribute.dtype)
    idx = att_logits.argmax(1, keepdim=True)
    est = attribute.clone()
    est.scatter_(1, idx, torch.maximum(est.gather(1, idx), none))
    return est


def swap_attributes(attribute: torch.Tensor, objs: torch.Tensor, obj_to_img: torch.Tensor, n_images: int,
                    matrix: torch.Tensor, rng):
    """train64.py:169-188 ("change GT attribute"), host side like the reference: in the first floor(N/3) images the first
    floor(n_obj/2) objects get 1-2 new attributes drawn from `matrix[obj]` (object-vs-attribute co-occurrence counts,
    matrix_obj_vs_att.pt) with the object's current attributes excluded.  rng: a `random.Random` (the reference draws
    randrange(1, 3) and then choices() from the global `random` module, in that order).  CPU tensors in, returns
    (swapped attribute, LongTensor of swapped rows); the step overwrites the same rows of attribute_est
    (train64.py:187-188) and keeps the ORIGINAL attribute as attribute_GT (train64.py:153, 241-244)."""
    out = attribute.clone()
    rows = []
    n_att = attribute.shape[1]
    for img_idx in range(n_images // 3):
        members = torch.nonzero(obj_to_img == img_idx).view(-1).tolist()
        for obj_idx in members[:len(members) // 2]:
            weights = matrix[int(objs[obj_idx])].clone()
            weights[torch.nonzero(attribute[obj_idx]).view(-1)] = 0
            k = rng.randrange(1, 3)
            new = rng.choices(range(n_att), weights, k=k)
            out[obj_idx] = 0
            out[obj_idx, torch.tensor(new, dtype=torch.long)] = 1
            rows.append(obj_idx)
    return out, torch.tensor(rows, dtype=torch.long)


class TrainStep:
    def __init__(self, image_size: int = 64, device="cuda", lr: float = 2e-4, lambdas: Optional[Dict[str, float]] = None,
                 pos_weight: Optional[torch.Tensor] = None, skip_dead_work: bool = True, fused_adam: bool = True,
                 capturable: bool = False, optimizer: str = "b200", att_matrix: Optional[torch.Tensor] = None,
                 swap_rng=None, fused_losses: bool = True):
        # fused_losses: the step arithmetic on the loss kernels of libb200gan (one launch per term, ops.FusedLoss) instead of
        # the PyTorch formulation kept below as `d_loss_torch` / `g_loss_torch` (same numbers; ~350 launches per iteration)
        self.fused_losses = fused_losses
        # att_matrix (179, 106) + swap_rng (random.Random): enable the reference's GT-attribute swap (train64.py:169-188)
        # in to_device(); without a matrix the batch's attributes are used as they are (the parity configuration)
        self.att_matrix, self.swap_rng = att_matrix, swap_rng
        self.image_size, self.obj_size = image_size, image_size // 2
        self.device = torch.device(device)
        self.lam = dict(LAMBDAS if lambdas is None else lambdas)
        self.netG, self.netD_image, self.netD_object, self.netD_att = [n.to(self.device) for n in
                                                                       build_networks(image_size)]
        self.pos_weight = (default_pos_weight() if pos_weight is None else pos_weight).to(self.device)
        self.fake_w = torch.tensor(FAKE_W, device=self.device)
        self.skip_dead_work = skip_dead_work
        kw = dict(lr=lr, betas=(0.5, 0.999))
        # optimizer: "b200" = the multi-tensor kernel (CUDA only) | "torch_fused" = torch.optim.Adam(fused=True) |
        # "torch" = torch.optim.Adam (on CUDA: fused=fused_adam).  Whichever updates the weights, the packed GEMM operands
        # follow through the global post-step hook (ops.WeightPacks).
        if self.device.type == "cuda" and optimizer == "b200":
            from .optim import Adam                                                  # multi-tensor kernel, device step counter
        else:
            Adam = torch.optim.Adam
            if optimizer == "torch_fused":
                kw.update(fused=True)
                if self.device.type == "cuda":
                    kw.update(capturable=capturable)
            elif self.device.type == "cuda":
                kw.update(fused=fused_adam, capturable=capturable)
        self.opt_G = Adam(self.netG.parameters(), **kw)                              # train64.py:111-114
        self.opt_D = [Adam(n.parameters(), **kw) for n in (self.netD_image, self.netD_object, self.netD_att)]
        self.d_nets = (self.netD_image, self.netD_object, self.netD_att)
        self.ddp_d = self.ddp_g = None
        self.sync_bn = False

    def enable_data_parallel(self, bucket_bytes: int = 25 << 20, group=None, sync_bn: bool = False):
        """Shard-local step + bucketed gradient all-reduce overlapped with backward (b200gan/ddp.py).

        sync_bn=False: DDP semantics — batch-norm statistics and loss means are per shard.
        sync_bn=True : exact global-batch semantics: every normalisation layer all-reduces its
        statistics (ops.set_sync_bn) and every loss mean is weighted by world * n_local / n_global (counts exchanged once per
        batch in to_device), so the averaged gradients equal those of the single-process step on the concatenated batch."""
        import torch.distributed as dist
        from .ddp import GradBucketer, broadcast_module
        for n in (self.netG,) + tuple(self.d_nets):
            broadcast_module(n, group=group)
        self.ddp_d = GradBucketer(self._d_bucket_order(), bucket_bytes, group)
        self.ddp_g = GradBucketer(self._g_bucket_order(), bucket_bytes, group)
        self.dp_group, self.sync_bn = group, bool(sync_bn)
        self.dp_world, self.dp_rank = dist.get_world_size(group), dist.get_rank(group)
        if sync_bn:
            if not self.fused_losses:
                raise ValueError("sync_bn needs the fused loss path (count-weighted terms)")
            ops.set_sync_bn(group if group is not None else True)

    def _g_bucket_order(self):
        """generator parameters in forward order of use (GradBucketer fills its buckets from the END of the list, i.e. in the
        order backward completes them): crop_encoder and attribute_encoder run first, then layout_encoder, global_encoder,
        decoder — the module registration order (generator_obj_att.py:609-616 puts the decoder in the middle and the
        attribute encoder last) would make the first bucket wait for the attribute encoder's gradients, the last to complete."""
        rank = {"crop_encoder": 0, "attribute_encoder": 1, "layout_encoder": 2, "global_encoder": 3, "decoder": 4}
        named = list(self.netG.named_parameters())
        named.sort(key=lambda kv: rank.get(kv[0].split(".")[0], 5))            # stable: registration order inside a module
        return [p for _, p in named]

    def _d_bucket_order(self):
        """discriminator parameters in the order their gradients complete, reversed (GradBucketer fills its buckets from the
        END of the list).  The three networks run their backward passes side by side (_side_by_side), so gradients of equal
        relative depth complete together: the parameters are merged by their fractional position inside their own network —
        with the plain concatenation every bucket would wait for the end of the slowest network's pass."""
        if not PARALLEL_D:
            return [p for n in self.d_nets for p in n.parameters()]
        keyed = []
        for ni, n in enumerate(self.d_nets):
            ps = list(n.parameters())
            for i, p in enumerate(ps):
                keyed.append(((i + 0.5) / len(ps), ni, i, p))
        keyed.sort(key=lambda t: t[:3])
        return [t[3] for t in keyed]

    def _global_counts(self, batch, b):
        """per-term weights world * n_local / n_global and the global image offset of this shard (sync_bn mode)"""
        import torch.distributed as dist
        N, O = batch["imgs"].shape[0], batch["objs"].shape[0]
        mine = torch.tensor([N, O, b["n_att_sel"], b.get("n_att_sel_g", b["n_att_sel"])], dtype=torch.float64, device=self.device)
        allc = [torch.zeros_like(mine) for _ in range(self.dp_world)]
        dist.all_gather(allc, mine, group=self.dp_group)
        allc = torch.stack(allc).cpu()
        tot = allc.sum(0)
        W = float(self.dp_world)
        n_tot, offset = int(tot[0]), int(allc[:self.dp_rank, 0].sum())
        n_change = math.floor(n_tot / 3)
        f = lambda i, mine_i: (W * mine_i / float(tot[i])) if float(tot[i]) > 0 else 0.0
        return dict(img=f(0, N), obj=f(1, O), att=f(2, b["n_att_sel"]), att_g=f(3, b.get("n_att_sel_g", b["n_att_sel"])), sum=W,
                    rec_mask=[0.0 if offset + i < n_change else 1.0 for i in range(N)], rec_denom=float(n_tot - n_change))

    # ---- batch handling ---------------------------------------------------------------------------------
    def to_device(self, batch: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        """train64.py:149-151: everything but obj_to_img moves to the device; index sets that the reference derives with
        nonzero() on the device are derived here from the CPU copy (no device sync inside the step).  With an att_matrix
        the GT-attribute swap (train64.py:169-188) is applied here, on the host as in the reference: `attribute` becomes
        the swapped one, `attribute_GT` keeps the original, `swap_rows` lists the changed objects.  A caller that did the
        swap itself passes `attribute_GT` (and optionally `swap_rows`) in the batch."""
        batch = dict(batch)
        if "attribute_GT" not in batch and self.att_matrix is not None:
            import random
            if self.swap_rng is None:
                self.swap_rng = random.Random()
            gt = batch["attribute"]
            batch["attribute"], batch["swap_rows"] = swap_attributes(gt, batch["objs"], batch["obj_to_img"],
                                                                     batch["imgs"].shape[0], self.att_matrix, self.swap_rng)
            batch["attribute_GT"] = gt
        b = {k: (v if k == "obj_to_img" else v.to(self.device, non_blocking=True)) for k, v in batch.items()}
        gt = batch.get("attribute_GT", batch["attribute"])
        b["att_idx"] = gt.sum(dim=1).nonzero().view(-1).to(self.device)                    # D-step: train64.py:241
        sel = (gt.sum(dim=1) != 0)
        b["att_sel"], b["n_att_sel"] = sel.float().to(self.device), int(sel.sum())           # the same set as a 0/1 row mask
        if "attribute_GT" in batch:
            b["att_idx_g"] = batch["attribute"].sum(dim=1).nonzero().view(-1).to(self.device)   # G-step: train64.py:323
            sel = (batch["attribute"].sum(dim=1) != 0)
            b["att_sel_g"], b["n_att_sel_g"] = sel.float().to(self.device), int(sel.sum())
        if self.sync_bn:
            b["dp"] = self._global_counts(batch, b)
        return b

    def generator(self, b, attribute_est, defer_tail=False):
        """defer_tail: see Generator.forward_batched — only _step passes it (and joins at the right places)"""
        return self.netG.forward_batched(b["imgs"], b["objs"], b["boxes"], b["masks"], b["obj_to_img"], b["z"],
                                         b["attribute"], b["masks_shift"], b["boxes_shift"], attribute_est,
                                         defer_tail=defer_tail)

    @staticmethod
    def _join(fake):
        """wait for the generator's deferred last call (Generator.forward_batched(defer_tail=True))"""
        j = fake.get("join")
        if j is not None:
            j()

    # ---- losses (train64.py:195-252, 284-364) ----------------------------------------------------------------
    # Calls of one discriminator on different inputs are batched along dim 0 as `groups` (in the reference's call order):
    # every spectral-normalised layer then runs `groups` power iterations and scales call g's rows by its own 1/sigma_g.
    def d_loss(self, b, fake):
        return self.d_loss_fused(b, fake) if self.fused_losses else self.d_loss_torch(b, fake)

    def g_loss(self, b, fake):
        return self.g_loss_fused(b, fake) if self.fused_losses else self.g_loss_torch(b, fake)

    def _side_by_side(self, thunks):
        """run independent sub-graphs (the three discriminators) on forked streams and join: their deep layers have fewer tiles
        than the GPU has CTA slots, so the networks fill the machine together.  Autograd runs every backward node on its
        forward stream, so the three backward passes overlap the same way.  Every fork starts by waiting for the current
        stream and every use of the results follows the join (memory handed between the streams' allocator pools is ordered
        by those two edges)."""
        if not (PARALLEL_D and self.device.type == "cuda" and ops.forks_enabled()) or len(thunks) < 2:
            return [t() for t in thunks]
        cur = torch.cuda.current_stream(self.device)
        side = ops.side_streams(self.device, len(thunks) - 1)
        out = [None] * len(thunks)
        for s in side:
            s.wait_stream(cur)
        out[0] = thunks[0]()
        for i, s in enumerate(side):
            with torch.cuda.stream(s):
                out[i + 1] = thunks[i + 1]()
        for s in side:
            cur.wait_stream(s)
        return out

    def d_loss_fused(self, b, fake):
        """train64.py:195-252 on the loss kernels: 4 term launches + 1 combine launch"""
        D_i, D_o, D_a = self.d_nets
        objs, lam = b["objs"], self.lam
        crops_input = fake["outputs"][0].detach()
        w4 = FAKE_W + (1.0,)
        dp = b.get("dp") or dict(img=1.0, obj=1.0, att=1.0, att_g=1.0, sum=1.0)     # count weights (sync_bn data parallel)
        acc = ops.FusedLoss(["d_img_fake", "d_img_real", "d_obj_fake", "d_obj_real", "d_obj_cls", "d_att"], self.device)
        src_i, (src, cls), att = self._side_by_side([
            lambda: D_i(torch.cat([fake["imgs_fake"].detach(), b["imgs"]]), groups=4),
            lambda: D_o(torch.cat([fake["crops_fake"].detach(), crops_input]), objs, groups=4),
            lambda: D_a(crops_input)])
        acc.add_bce_groups("d_img_fake", src_i, 4, (0, 0, 0, 1), w4, lam["img_adv"] * dp["img"], split_group=3)
        acc.add_bce_groups("d_obj_fake", src, 4, (0, 0, 0, 1), w4, lam["obj_adv"] * dp["obj"], split_group=3)
        acc.add_ce_groups("d_obj_cls", cls, objs, 4, (0, 0, 0, 1), lam["obj_cls"] * dp["obj"])
        acc.add_bce_pos_weight_rows("d_att", att, b["attribute_GT"], b["att_sel"], b["n_att_sel"], self.pos_weight,
                                    1, (1.0,), lam["att_cls"] * dp["att"])
        total = acc.total()
        return total, acc.term_dict(dict(d_img_fake=lam["img_adv"], d_img_real=lam["img_adv"], d_obj_fake=lam["obj_adv"],
                                         d_obj_real=lam["obj_adv"], d_obj_cls=lam["obj_cls"], d_att=lam["att_cls"]))

    def g_loss_fused(self, b, fake):
        """train64.py:284-364 on the loss kernels: 7 term launches + 1 combine launch"""
        (crops_input, crops_input_rec, crops_rand, crops_shift, img_rec, img_rand, img_shift, mu, logvar, z_rand_rec,
         z_rand_shift) = fake["outputs"]
        D_i, D_o, D_a = self.d_nets
        imgs, z, objs, lam = b["imgs"], b["z"], b["objs"], self.lam
        N = imgs.shape[0]
        n_change = math.floor(N / 3)
        dp = b.get("dp")
        if dp is None:      # single process / DDP semantics: the first floor(N/3) images of THIS batch are excluded (train64.py:284-287)
            dp = dict(img=1.0, obj=1.0, att=1.0, att_g=1.0, sum=1.0, rec_mask=[0.0] * n_change + [1.0] * (N - n_change),
                      rec_denom=float(N - n_change))
        rec_mask = ops._loss_const(dp["rec_mask"], self.device)
        acc = ops.FusedLoss(["g_img_rec", "g_z_rec", "g_kl", "g_img_adv", "g_obj_adv", "g_obj_cls", "g_obj_att"], self.device)
        acc.add_l1_rows("g_img_rec", img_rec, imgs, N, rec_mask, dp["rec_denom"], lam["img_rec"] * dp["sum"])
        acc.add_kl("g_kl", mu, logvar, lam["kl"] * dp["sum"])
        src_i, (src, cls), att = self._side_by_side([
            lambda: D_i(fake["imgs_fake"], groups=3),
            lambda: D_o(fake["crops_fake"], objs, groups=3),
            lambda: D_a(fake["crops_fake"], groups=3)])
        self._join(fake)
        # 0.5 * mean|z_rand_rec - z| + 0.5 * mean|z_rand_shift - z|: the two crop-encoder passes are rows of one (2, O*z) tensor
        acc.add_l1_rows("g_z_rec", fake["mu2"], z, 2, None, 1.0, 0.5 * lam["z_rec"] * dp["obj"], broadcast_b=True)
        acc.add_bce_groups("g_img_adv", src_i, 3, (1, 1, 1), FAKE_W, lam["img_adv"] * dp["img"])
        acc.add_bce_groups("g_obj_adv", src, 3, (1, 1, 1), FAKE_W, lam["obj_adv"] * dp["obj"])
        acc.add_ce_groups("g_obj_cls", cls, objs, 3, FAKE_W, lam["obj_cls"] * dp["obj"])
        acc.add_bce_pos_weight_rows("g_obj_att", att, b["attribute"], b.get("att_sel_g", b["att_sel"]),
                                    b.get("n_att_sel_g", b["n_att_sel"]), self.pos_weight, 3, FAKE_W, lam["att_cls"] * dp["att_g"])
        total = acc.total()
        return total, acc.term_dict(dict(g_img_rec=lam["img_rec"], g_z_rec=lam["z_rec"], g_kl=lam["kl"], g_img_adv=lam["img_adv"],
                                         g_obj_adv=lam["obj_adv"], g_obj_cls=lam["obj_cls"], g_obj_att=lam["att_cls"]))

    def d_loss_torch(self, b, fake):
        D_i, D_o, D_a = self.d_nets
        objs = b["objs"]
        N, O = b["imgs"].shape[0], objs.shape[0]
        crops_input = fake["outputs"][0].detach()
        w = self.fake_w
        l = {}
        # D_image on img_rec, img_rand, img_shift (fake, detached), then the real images   (train64.py:195-212)
        src = D_i(torch.cat([fake["imgs_fake"].detach(), b["imgs"]]), groups=4)
        l["d_img_fake"] = (w * bce_const_groups(src[:3 * N], 0, 3)).sum()
        l["d_img_real"] = bce_const(src[3 * N:], 1)
        # D_object on crops_input_rec, crops_rand, crops_shift, then the real crops   (train64.py:215-238)
        src, cls = D_o(torch.cat([fake["crops_fake"].detach(), crops_input]), objs, groups=4)
        l["d_obj_fake"] = (w * bce_const_groups(src[:3 * O], 0, 3)).sum()
        l["d_obj_real"] = bce_const(src[3 * O:], 1)
        l["d_obj_cls"] = F.cross_entropy(cls[3 * O:], objs)
        att_cls = D_a(crops_input)
        idx = b["att_idx"]
        l["d_att"] = F.binary_cross_entropy_with_logits(att_cls.index_select(0, idx), b["attribute_GT"].index_select(0, idx),
                                                        pos_weight=self.pos_weight)
        lam = self.lam
        total = lam["img_adv"] * (l["d_img_fake"] + l["d_img_real"]) + lam["obj_adv"] * (l["d_obj_fake"] + l["d_obj_real"]) \
            + lam["obj_cls"] * l["d_obj_cls"] + lam["att_cls"] * l["d_att"]
        return total, l

    def g_loss_torch(self, b, fake):
        (crops_input, crops_input_rec, crops_rand, crops_shift, img_rec, img_rand, img_shift, mu, logvar, z_rand_rec,
         z_rand_shift) = fake["outputs"]
        self._join(fake)
        D_i, D_o, D_a = self.d_nets
        imgs, z, objs, attribute = b["imgs"], b["z"], b["objs"], b["attribute"]
        N, O = imgs.shape[0], objs.shape[0]
        n_change = math.floor(N / 3)
        rec_mask = torch.ones(N, device=imgs.device)
        rec_mask[:n_change] = 0
        w = self.fake_w
        l = {}
        l["g_img_rec"] = (rec_mask * (img_rec - imgs).abs().view(N, -1).mean(1)).sum() / (N - n_change)
        l["g_z_rec"] = 0.5 * (z_rand_rec - z).abs().mean() + 0.5 * (z_rand_shift - z).abs().mean()
        l["g_kl"] = -0.5 * torch.sum(1 + logvar - mu.pow(2) - logvar.exp())
        l["g_img_adv"] = (w * bce_const_groups(D_i(fake["imgs_fake"], groups=3), 1, 3)).sum()
        # D_object / D_att on crops_input_rec, crops_rand, crops_shift   (train64.py:298-349)
        src, cls = D_o(fake["crops_fake"], objs, groups=3)
        att = D_a(fake["crops_fake"], groups=3)
        idx = b.get("att_idx_g", b["att_idx"])
        n_idx = idx.numel()
        idx3 = torch.cat([idx, idx + O, idx + 2 * O])
        att_t = attribute.index_select(0, idx)
        l["g_obj_adv"] = (w * bce_const_groups(src, 1, 3)).sum()
        l["g_obj_cls"] = (w * F.cross_entropy(cls, objs.repeat(3), reduction="none").view(3, O).mean(1)).sum()
        bce = F.binary_cross_entropy_with_logits(att.index_select(0, idx3), att_t.repeat(3, 1), pos_weight=self.pos_weight,
                                                 reduction="none")
        l["g_obj_att"] = (w * bce.view(3, n_idx * bce.shape[1]).mean(1)).sum()
        lam = self.lam
        total = lam["img_rec"] * l["g_img_rec"] + lam["z_rec"] * l["g_z_rec"] + lam["img_adv"] * l["g_img_adv"] \
            + lam["obj_adv"] * l["g_obj_adv"] + lam["obj_cls"] * l["g_obj_cls"] + lam["att_cls"] * l["g_obj_att"] \
            + lam["kl"] * l["g_kl"]
        return total, l

    # ---- the step --------------------------------------------------------------------------------------------
    def step(self, b: Dict[str, torch.Tensor], optimizer_step: bool = True, seeds=None):
        """b: device batch from to_device().  seeds: optional (seed_d, seed_g) for torch.manual_seed before each generator
        forward (pins the CropEncoder noise like the parity harness of the oracle)."""
        from . import nn as bnn
        with bnn.deferred_batch_counts():
            return self._step(b, optimizer_step, seeds)

    def _step(self, b, optimizer_step, seeds):
        from models.bilinear import crop_bbox_batch
        D_i, D_o, D_a = self.d_nets
        b = dict(b)
        if "attribute_GT" not in b:
            b["attribute_GT"] = b["attribute"]                                                  # no swap: train64.py:153
        with torch.no_grad():
            crops = crop_bbox_batch(b["imgs"], b["boxes"], b["obj_to_img"], self.obj_size)     # train64.py:160
            est_logits = D_a(crops)                                                             # train64.py:161
        attribute_est = estimate_attributes(est_logits, b["attribute_GT"])
        rows = b.get("swap_rows")
        if rows is not None and rows.numel() > 0:                                               # train64.py:187-188
            attribute_est.index_copy_(0, rows, b["attribute"].index_select(0, rows))
        if DEFER_TAIL and self.device.type == "cuda" and ops.forks_enabled():
            gen = lambda: self.generator(b, attribute_est, True)          # noqa: E731
        else:
            gen = lambda: self.generator(b, attribute_est)                # noqa: E731
        # ---------------- D-step ----------------
        if seeds is not None:
            torch.manual_seed(seeds[0])
        if self.skip_dead_work:
            with torch.no_grad():
                fake = gen()                                                                   # train64.py:191
        else:
            fake = gen()
        # The G-step's generator pass (train64.py:280) reads nothing the D-step writes (discriminator weights, their
        # spectral-norm vectors): it is issued HERE on a forked stream, after the D-step pass it must follow (batch-norm
        # running statistics, noise draws), and runs next to the discriminators' forward / backward / Adam; joined before the
        # G-step losses.  Its latency-bound parts (ConvLSTM time steps, 8x8 / 16x16 layers) fill the gaps of the D-step.
        out = s2 = None
        if OVERLAP_G2 and self.device.type == "cuda" and ops.forks_enabled():
            cur = torch.cuda.current_stream(self.device)
            s2 = ops.side_streams(self.device, 4)[3]
            s2.wait_stream(cur)
            if seeds is not None:
                torch.manual_seed(seeds[1])
            with torch.cuda.stream(s2):
                self._join(fake)                        # (the D-step pass's deferred crop_encoder call comes first)
                out = gen()
        d_total, d_terms = self.d_loss(b, fake)
        for n in self.d_nets:
            n.zero_grad(set_to_none=True)
        if self.ddp_d is not None:
            self.ddp_d.arm()
        d_total.backward()
        if self.ddp_d is not None:
            self.ddp_d.finish()
        self._join(fake)              # the D-step generator pass's deferred call (its running-statistics updates) ends here
        if optimizer_step:
            for o in self.opt_D:
                o.step()
        # ---------------- G-step ----------------
        if s2 is not None:
            torch.cuda.current_stream(self.device).wait_stream(s2)
        elif seeds is not None:
            torch.manual_seed(seeds[1])
        if self.skip_dead_work:
            for n in self.d_nets:
                for p in n.parameters():
                    p.requires_grad_(False)
        try:
            if out is None:
                out = gen()                                                                       # train64.py:280
            g_total, g_terms = self.g_loss(b, out)
            self.netG.zero_grad(set_to_none=True)
            if self.ddp_g is not None:
                self.ddp_g.arm()
            g_total.backward()
            if self.ddp_g is not None:
                self.ddp_g.finish()
        finally:
            if self.skip_dead_work:
                for n in self.d_nets:
                    for p in n.parameters():
                        p.requires_grad_(True)
        if optimizer_step:
            self.opt_G.step()
        return dict(d_loss=d_total.detach(), g_loss=g_total.detach(), d_terms={k: v.detach() for k, v in d_terms.items()},
                    g_terms={k: v.detach() for k, v in g_terms.items()}, out_g=[t.detach() for t in out["outputs"]],
                    attribute_est=attribute_est)
