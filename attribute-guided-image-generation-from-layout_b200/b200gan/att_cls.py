"""Stand-alone attribute-classifier training iteration (evaluation/train_att_cls.py:196-258): the
spectrally-normalised AttributeDiscriminator trained on real object crops with the pos-weighted BCE of the annotated objects —
the classifier `test64.py:103` loads to score attribute edits.  Same modules, crop kernel, loss kernel and optimizer as the G+D
step; one call = crop -> classify -> loss -> backward -> Adam (+ GEMM-operand re-pack through the optimizer hook)."""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import ops
from .step import NUM_ATTRIBUTES, default_pos_weight


class AttributeClassifierStep:
    def __init__(self, crop_size: int = 64, device="cuda", lr: float = 2e-4, pos_weight: Optional[torch.Tensor] = None,
                 optimizer: str = "b200"):
        from models.bilinear import crop_bbox_batch  # noqa: F401  (the reference's entry point, resolved at call time)
        from models.discriminator import AttributeDiscriminator, add_sn
        self.crop_size, self.device = crop_size, torch.device(device)
        self.net = add_sn(AttributeDiscriminator(n_attribute=NUM_ATTRIBUTES)).to(self.device)       # train_att_cls.py:199-200
        self.pos_weight = (default_pos_weight() if pos_weight is None else pos_weight).to(self.device)
        if self.device.type == "cuda" and optimizer == "b200":
            from .optim import Adam
        else:
            Adam = torch.optim.Adam
        self.opt = Adam(self.net.parameters(), lr=lr, betas=(0.5, 0.999))                               # train_att_cls.py:202

    def to_device(self, batch: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        b = {k: (v if k == "obj_to_img" else v.to(self.device, non_blocking=True)) for k, v in batch.items()
             if k in ("imgs", "boxes", "obj_to_img", "attribute")}
        sel = batch["attribute"].sum(dim=1) != 0                                                      # train_att_cls.py:233
        b["att_sel"], b["n_att_sel"] = sel.float().to(self.device), int(sel.sum())
        return b

    def step(self, b: Dict[str, torch.Tensor], optimizer_step: bool = True):
        from models.bilinear import crop_bbox_batch
        with torch.no_grad():
            crops = crop_bbox_batch(b["imgs"], b["boxes"], b["obj_to_img"], self.crop_size)          # train_att_cls.py:230
        logits = self.net(crops)                                                                       # :232
        acc = ops.FusedLoss(["att_cls"], self.device)
        acc.add_bce_pos_weight_rows("att_cls", logits, b["attribute"], b["att_sel"], b["n_att_sel"], self.pos_weight, 1,
                                    (1.0,), 1.0)                                                       # :233-237
        loss = acc.total()
        self.net.zero_grad(set_to_none=True)
        loss.backward()
        if optimizer_step:
            self.opt.step()
        return dict(loss=loss.detach(), logits=logits.detach())
