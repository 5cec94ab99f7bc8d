"""-m gpu: the box -> layout integer work (bit-exact), BASELINE config 4 (eval-mode generator forward, test64.py style) and
config 5 (layout microbench shapes: rasterise + broadcast + crops at 128x128 with 8..30 boxes per image)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from b200gan import _lib, layout, ops  # noqa: E402
from b200gan.step import TrainStep  # noqa: E402
from helpers import load_states, rel  # noqa: E402
from oracle import gan_oracle as O  # noqa: E402


def _edge_boxes(seed=0, n=300):
    g = torch.Generator().manual_seed(seed)
    xy0 = torch.rand(n, 2, generator=g) * 0.7
    wh = torch.rand(n, 2, generator=g) * 0.5
    boxes = torch.cat([xy0, (xy0 + wh).clamp(max=1.0)], 1)
    # products that land exactly on .5 (Python round is half-to-even): k/128 * 64 = k/2
    k = torch.randint(0, 129, (n // 3, 4), generator=g).float() / 128.0
    k = torch.cat([torch.minimum(k[:, :2], k[:, 2:]), torch.maximum(k[:, :2], k[:, 2:])], 1)
    special = torch.tensor([[0.0, 0.0, 1.0, 1.0], [0.5, 0.5, 0.5, 0.5], [0.9, 0.9, 0.1, 0.1], [0.2578125, 0.0, 0.5078125, 1.0],
                            [0.0078125, 0.0234375, 0.0390625, 0.0546875], [-0.05, 0.1, 0.4, 1.2], [0.3, -0.2, 1.3, 0.6],
                            [0.49, 0.0, 0.51, 1.0], [0.0, 0.3, 0.49999, 0.31]])
    return torch.cat([boxes, k, special]).float().contiguous()


@pytest.mark.parametrize("H,W", [(64, 64), (128, 128), (66, 30), (7, 5)])
def test_rasterize_boxes_bit_exact(H, W):
    """round-half-even on double + Python slice semantics, against the loader restatement (oracle), bit for bit"""
    boxes = _edge_boxes(H)
    got = layout.rasterize_boxes(boxes.cuda(), H, W)
    want = O.rasterize_boxes(boxes, H, W)
    assert got.shape == want.shape and torch.equal(got.cpu(), want)


def test_shift_boxes_bit_exact():
    boxes = _edge_boxes(3)
    got = layout.shift_boxes(boxes.cuda())
    want = O.shift_boxes(boxes)
    assert torch.equal(got.cpu(), want)
    m, bs, ms = layout.layout_inputs(boxes.cuda(), 64)
    assert torch.equal(ms.cpu(), O.rasterize_boxes(want, 64, 64)) and torch.equal(bs.cpu(), want)


def test_rasterize_empty_and_requires_cuda():
    assert layout.rasterize_boxes(torch.zeros(0, 4).cuda(), 64, 64).shape == (0, 1, 64, 64)
    with pytest.raises(_lib.B200Error):
        layout.rasterize_boxes(torch.zeros(2, 4), 64, 64)


# ---------------------------------------------------------------------------------------------------------
# config 5: layout microbench shapes at 128x128
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("per_image", [8, 12, 16, 20, 24, 30])
def test_config5_layout_and_crops_128(per_image):
    """128x128, `per_image` boxes per image: device rasteriser == loader masks; the embedding (x) mask broadcast
    (LayoutEncoder c0 as rank-1 outer product) and the 64x64 crops forward / backward against the oracle."""
    N, H, S = 4, 128, 64
    b = O.synth_batch(N, H, per_image, seed=per_image)
    boxes, o2i = b["boxes"], b["obj_to_img"]
    On = boxes.shape[0]
    masks = layout.rasterize_boxes(boxes.cuda(), H, H)
    assert torch.equal(masks.cpu(), b["masks"])
    # broadcast: v (O, C) (x) mask with a zero ring  ==  (v.view(O,C,1,1) * mask) zero-padded by 1   (generator_obj_att.py:489-490, c0 pad 1)
    g = torch.Generator().manual_seed(1)
    v = torch.randn(On, 64, generator=g)
    out = _lib.K.mask_outer_fwd(v.cuda(), masks, On, H, H, 64)
    want = torch.nn.functional.pad(v.view(On, 64, 1, 1) * b["masks"], (1, 1, 1, 1)).permute(0, 2, 3, 1)
    assert torch.equal(out.cpu(), want.contiguous())
    # crops: forward values, bit-exact tap indices, deterministic backward
    feats = torch.randn(N, 3, H, H, generator=g)
    fr = feats.clone().requires_grad_(True)
    want_c = O.crop_bbox_batch(fr, boxes, o2i, S)
    fd = feats.cuda().requires_grad_(True)
    got_c = ops.crop_bbox_batch(fd, boxes.cuda(), o2i, S)
    assert float((got_c.detach().cpu() - want_c.detach()).abs().max()) < 2e-6
    wgt = torch.randn(want_c.shape, generator=g)
    (want_c * wgt).sum().backward()
    (got_c * wgt.cuda()).sum().backward()
    assert rel(fd.grad, fr.grad) < 1e-5
    ix0, iy0, _, _ = _lib.K.crop_taps(boxes.cuda(), ops.crop_weights(S, "cuda"), ops.crop_weights(S, "cuda"), H, H, S, S)
    rx0, ry0, _, _ = O.crop_taps(boxes, H, H, S, S)
    assert torch.equal(ix0.cpu(), rx0) and torch.equal(iy0.cpu(), ry0)


# ---------------------------------------------------------------------------------------------------------
# config 4: eval-mode generator forward (test64.py:114-198 style)
# ---------------------------------------------------------------------------------------------------------
def test_config4_eval_generator_matches_oracle():
    """netG.eval(): running-statistics BN / CBN / SPADE, no running-stat updates; all 11 outputs against the oracle in
    eval mode with the same CropEncoder noise (fp32 kernels, 1e-4)."""
    ops.set_precision("fp32")
    states = O.make_states(64, 2)
    batch = O.synth_batch(3, 64, None, 21)
    ts = TrainStep(64, device="cuda")
    load_states(ts, states)
    ts.netG.eval()
    before = {k: v.clone() for k, v in ts.netG.state_dict().items() if "running" in k or "num_batches" in k}
    b = ts.to_device(batch)
    torch.manual_seed(77)
    with torch.no_grad():
        got = ts.netG(b["imgs"], b["objs"], b["boxes"], b["masks"], b["obj_to_img"], b["z"], b["attribute"], b["masks_shift"],
                      b["boxes_shift"], b["attribute"])
    torch.manual_seed(77)
    with torch.no_grad():
        want = O.generator_forward(states["G"], batch["imgs"], batch["objs"], batch["boxes"], batch["masks"],
                                   batch["obj_to_img"], batch["z"], batch["attribute"], batch["masks_shift"],
                                   batch["boxes_shift"], batch["attribute"], training=False, image_size=64)
    for i, (a, r) in enumerate(zip(got, want)):
        assert rel(a, r) < 1e-4, "eval-mode generator output %d: %.3e" % (i, rel(a, r))
    after = ts.netG.state_dict()
    for k, v in before.items():
        assert torch.equal(after[k], v), "eval mode must not touch %s" % k


def test_config4_full_size_properties():
    """BASELINE config 4 shape: batch 256, 8 objects/image, generator-only eval forward in bf16 — finite images, generated
    crops equal crops of the generated images, and images are independent of the rest of the batch (eval mode)."""
    ops.set_precision("bf16")
    try:
        ts = TrainStep(64, device="cuda")
        ts.netG.eval()
        ts.netG.crop_encoder.eps_source = lambda o, z, d: torch.zeros(o, z, device=d)
        batch = O.synth_batch(256, 64, 8, 5)
        b = ts.to_device(batch)
        with torch.no_grad():
            out = ts.netG(b["imgs"], b["objs"], b["boxes"], b["masks"], b["obj_to_img"], b["z"], b["attribute"],
                          b["masks_shift"], b["boxes_shift"], b["attribute"])
            assert out[5].shape == (256, 3, 64, 64) and all(torch.isfinite(t).all() for t in out)
            again = ops.crop_bbox_batch(out[5], b["boxes"], b["obj_to_img"], 32)
            assert torch.equal(again, out[2])
            sub = {k: (v[:64 * 8] if v.shape[0] == 2048 else v[:64]) for k, v in batch.items()}
            bs = ts.to_device(sub)
            part = ts.netG(bs["imgs"], bs["objs"], bs["boxes"], bs["masks"], bs["obj_to_img"], bs["z"], bs["attribute"],
                           bs["masks_shift"], bs["boxes_shift"], bs["attribute"])
            assert rel(part[5], out[5][:64]) < 2e-2       # bf16 tiles differ with the batch size; eval BN is per sample
    finally:
        ops.set_precision("fp32")


# ---------------------------------------------------------------------------------------------------------
# multi-tensor Adam
# ---------------------------------------------------------------------------------------------------------
def test_multi_tensor_adam_matches_torch():
    """b200gan.optim.Adam (one launch, device step counter) against torch.optim.Adam(lr=2e-4, betas=(0.5, 0.999)) over
    several steps on tensors of odd sizes (vector and tail paths), incl. inside a CUDA graph"""
    from b200gan.optim import Adam
    g = torch.Generator().manual_seed(0)
    shapes = [(1024, 512, 3, 3), (179,), (7, 5), (64, 3, 7, 7), (1,), (65537,), (256, 257)]
    ref = [torch.randn(s, generator=g).cuda().requires_grad_(True) for s in shapes]
    mine = [p.detach().clone().requires_grad_(True) for p in ref]
    o_ref = torch.optim.Adam(ref, lr=2e-4, betas=(0.5, 0.999))
    o_b = Adam(mine, lr=2e-4, betas=(0.5, 0.999))
    for it in range(4):
        for a, b in zip(ref, mine):
            gr = torch.randn(a.shape, generator=g).cuda() * (10.0 ** (it - 2))
            a.grad, b.grad = gr.clone(), gr.clone()
        o_ref.step()
        o_b.step()
    for a, b in zip(ref, mine):
        assert rel(b, a) < 1e-6
        assert rel(o_b.state[b]["exp_avg_sq"], o_ref.state[a]["exp_avg_sq"]) < 1e-6
    assert float(o_b.state[mine[0]]["step"]) == 4.0
    # capturable: the same update replayed from a CUDA graph keeps counting on the device
    for b in mine:
        b.grad = torch.ones_like(b)
    gph = torch.cuda.CUDAGraph()
    o_b.step()
    torch.cuda.synchronize()
    with torch.cuda.graph(gph):
        o_b.step()
    gph.replay()
    torch.cuda.synchronize()
    assert float(o_b.state[mine[0]]["step"]) == 6.0      # 4 + 1 eager + 1 replay (capture itself does not run)
    sd = o_b.state_dict()
    assert set(sd["state"][0].keys()) == {"step", "exp_avg", "exp_avg_sq"}


def test_adam_save_load_step_parity():
    """checkpoint round trip (utils/model_saver_iter.py:40-57 saves optimizer.state_dict()): a b200gan Adam that has already
    stepped loads another optimizer's state and must continue exactly like torch.optim.Adam does after the same load —
    bias-correction step count and moments come from the checkpoint, not from the instance's own history."""
    import copy
    from b200gan.optim import Adam
    g = torch.Generator().manual_seed(1)
    shapes = [(300, 7), (65537,), (64, 3, 3, 3)]
    ref = [torch.randn(s, generator=g).cuda().requires_grad_(True) for s in shapes]
    mine = [p.detach().clone().requires_grad_(True) for p in ref]
    o_ref = torch.optim.Adam(ref, lr=2e-4, betas=(0.5, 0.999))
    o_b = Adam(mine, lr=2e-4, betas=(0.5, 0.999))

    def both(n, scale=1.0):
        for _ in range(n):
            for a, b in zip(ref, mine):
                gr = torch.randn(a.shape, generator=g).cuda() * scale
                a.grad, b.grad = gr.clone(), gr.clone()
            o_ref.step()
            o_b.step()

    both(3)
    ckpt_ref, ckpt_b = copy.deepcopy(o_ref.state_dict()), copy.deepcopy(o_b.state_dict())
    assert float(ckpt_b["state"][0]["step"]) == 3.0
    both(2, 5.0)                                           # the instances move on ...
    with torch.no_grad():
        for a, b in zip(ref, mine):                        # ... then weights and optimizer state are restored
            b.copy_(a)
    o_ref.load_state_dict(ckpt_ref)
    o_b.load_state_dict(ckpt_b)
    assert float(o_b.state[mine[0]]["step"]) == 3.0
    both(2)
    for a, b in zip(ref, mine):
        assert rel(b, a) < 1e-6
        assert rel(o_b.state[b]["exp_avg"], o_ref.state[a]["exp_avg"]) < 1e-6
        assert rel(o_b.state[b]["exp_avg_sq"], o_ref.state[a]["exp_avg_sq"]) < 1e-6
    assert float(o_b.state[mine[0]]["step"]) == 5.0
    # loading a torch.optim.Adam checkpoint into the kernel optimizer works too (same state layout)
    o_b.load_state_dict(copy.deepcopy(o_ref.state_dict()))
    both(1)
    for a, b in zip(ref, mine):
        assert rel(b, a) < 1e-6


# ---------------------------------------------------------------------------------------------------------
# masks_to_layout (config 5; utils/draw_box.py:482-483) against oracle/layout_oracle.py
# ---------------------------------------------------------------------------------------------------------
def _m2l_inputs(seed, N, per_image, D, M, ragged=False):
    from oracle import layout_oracle as LO  # noqa: F401
    g = torch.Generator().manual_seed(seed)
    counts = [int(torch.randint(0, per_image + 1, (1,), generator=g)) if ragged else per_image for _ in range(N)]
    if ragged:
        counts[0] = 0                       # an image without objects stays all zero
        counts[-1] = max(counts[-1], 33)    # more than one 32-object pass
    O_ = sum(counts)
    o2i = torch.cat([torch.full((c,), i, dtype=torch.long) for i, c in enumerate(counts)]) if O_ else torch.zeros(0, dtype=torch.long)
    xy0 = torch.rand(O_, 2, generator=g) * 0.7
    wh = torch.rand(O_, 2, generator=g) * 0.5 + 0.05
    boxes = torch.cat([xy0, (xy0 + wh).clamp(max=1.0)], 1)
    if O_ > 3:
        boxes[0] = torch.tensor([0.0, 0.0, 1.0, 1.0])
        boxes[1] = torch.tensor([0.25, 0.5, 0.75, 1.0])
        boxes[2] = torch.tensor([-0.1, 0.2, 0.4, 1.3])          # partly outside the canvas
    vecs = torch.randn(O_, D, generator=g)
    masks = torch.rand(O_, M, M, generator=g)
    return vecs, boxes, masks, o2i


@pytest.mark.parametrize("N,per_image,D,M,H,ragged", [(4, 8, 128, 16, 128, False), (3, 30, 64, 16, 64, False),
                                                      (5, 12, 8, 7, 33, True), (2, 5, 4, 1, 16, False)])
def test_masks_to_layout_matches_restatement(N, per_image, D, M, H, ragged):
    from oracle import layout_oracle as LO
    vecs, boxes, masks, o2i = _m2l_inputs(N * 7 + D, N, per_image, D, M, ragged)
    W = H if H != 33 else 20
    # index work: bit-exact
    ix0, iy0, fx, fy = layout.masks_to_layout_taps(boxes.cuda(), M, H, W)
    rx0, ry0, rfx, rfy = LO.layout_taps(boxes, M, H, W)
    assert torch.equal(ix0.cpu(), rx0) and torch.equal(iy0.cpu(), ry0)
    assert torch.equal(fx.cpu(), rfx) and torch.equal(fy.cpu(), rfy)
    # forward + gradients w.r.t. the embeddings and the masks
    v, m = vecs.clone().requires_grad_(True), masks.clone().requires_grad_(True)
    want = LO.masks_to_layout(v, boxes, m, o2i, H, W, N=N)
    gw = torch.randn(want.shape, generator=torch.Generator().manual_seed(1))
    (want * gw).sum().backward()
    vc, mc = vecs.cuda().requires_grad_(True), masks.cuda().requires_grad_(True)
    got = layout.masks_to_layout(vc, boxes.cuda(), mc, o2i, H, W, N=N)
    assert got.shape == want.shape
    assert float((got.cpu() - want.detach()).abs().max()) <= 1e-5 * max(1.0, float(want.abs().max()))
    (got * gw.cuda()).sum().backward()
    assert rel(vc.grad, v.grad) < 1e-5 and rel(mc.grad, m.grad) < 1e-5
    # deterministic: a second evaluation is bit-identical (gather, fixed-order reductions)
    vc2, mc2 = vecs.cuda().requires_grad_(True), masks.cuda().requires_grad_(True)
    got2 = layout.masks_to_layout(vc2, boxes.cuda(), mc2, o2i, H, W, N=N)
    (got2 * gw.cuda()).sum().backward()
    assert torch.equal(got2, got) and torch.equal(vc2.grad, vc.grad) and torch.equal(mc2.grad, mc.grad)


def test_masks_to_layout_full_size_properties():
    """config 5 at full size (128x128, N = 32, 30 boxes per image, D = 128): the oracle would need a 32 GB (O,D,H,W)
    intermediate, so size-independent properties: linearity in the embeddings, per-image independence, a one-hot embedding
    reproduces the resampled mask (draw_box.py:474-484 `cropped_to_full_mask`), unsorted obj_to_img == sorted."""
    from oracle import layout_oracle as LO
    N, per, D, M, H = 32, 30, 128, 16, 128
    vecs, boxes, masks, o2i = _m2l_inputs(3, N, per, D, M)
    bc, mc = boxes.cuda(), masks.cuda()
    a = layout.masks_to_layout(vecs.cuda(), bc, mc, o2i, H, N=N)
    assert torch.isfinite(a).all()
    v2 = torch.randn(vecs.shape, generator=torch.Generator().manual_seed(9))
    b = layout.masks_to_layout(v2.cuda(), bc, mc, o2i, H, N=N)
    ab = layout.masks_to_layout((2.0 * vecs - 0.5 * v2).cuda(), bc, mc, o2i, H, N=N)
    assert rel(ab, 2.0 * a - 0.5 * b) < 1e-5
    sel = o2i < 2                                                # the first two images alone give the same planes
    part = layout.masks_to_layout(vecs[sel].cuda(), boxes[sel].cuda(), masks[sel].cuda(), o2i[sel], H, N=2)
    assert torch.equal(part, a[:2])
    want = LO.masks_to_layout(vecs[sel], boxes[sel], masks[sel], o2i[sel], H, N=2)
    assert float((part.cpu() - want).abs().max()) < 1e-4
    perm = torch.randperm(o2i.numel(), generator=torch.Generator().manual_seed(2))
    shuf = layout.masks_to_layout(vecs[perm].cuda(), boxes[perm].cuda(), masks[perm].cuda(), o2i[perm], H, N=N)
    assert rel(shuf, a) < 1e-6
    eye = torch.eye(per)[:, :per]                                # one-hot "embedding" per object of image 0
    k = int((o2i == 0).sum())
    full = layout.masks_to_layout(torch.eye(k, 32)[:, :32].contiguous().cuda(), boxes[:k].cuda(), masks[:k].cuda(),
                                  torch.zeros(k, dtype=torch.long), H, N=1)
    ref = LO.masks_to_layout(torch.eye(k, 32), boxes[:k], masks[:k], torch.zeros(k, dtype=torch.long), H, N=1)
    assert float((full.cpu() - ref).abs().max()) < 1e-5


# ---------------------------------------------------------------------------------------------------------
# data contract on the device: collate + one-hot attributes + imagenet_deprocess_batch
# ---------------------------------------------------------------------------------------------------------
def test_imagenet_deprocess_bit_exact():
    import os
    from b200gan import data
    from helpers import GOLD
    from oracle import data_oracle as DO
    g = torch.load(os.path.join(GOLD, "data.pt"))                      # produced by the unmodified data/utils.py
    assert torch.equal(data.imagenet_deprocess_batch(g["imgs"].cuda(), True).cpu(), g["out_rescale"])
    assert torch.equal(data.imagenet_deprocess_batch(g["imgs"].cuda(), False).cpu(), g["out_plain"])
    big = torch.randn(32, 3, 128, 128, generator=torch.Generator().manual_seed(5)) * 2.0
    got = data.imagenet_deprocess_batch(big.cuda())
    assert got.dtype == torch.uint8 and torch.equal(got.cpu(), DO.imagenet_deprocess_batch(big))
    assert int(got.amin()) == 0 and int(got.amax()) == 255                # every image spans the full range after rescale


def test_collate_on_device_matches_loader_restatement():
    from b200gan import data
    from oracle import data_oracle as DO
    g = torch.Generator().manual_seed(8)
    samples = []
    for i in range(6):
        n = int(torch.randint(3, 10, (1,), generator=g))
        xy0 = torch.rand(n, 2, generator=g) * 0.6
        boxes = torch.cat([xy0, (xy0 + torch.rand(n, 2, generator=g) * 0.4 + 0.05).clamp(max=1.0)], 1)
        att = torch.full((n, 30), -1, dtype=torch.long)
        for r in range(n):
            k = int(torch.randint(0, 4, (1,), generator=g))
            att[r, :k] = torch.randperm(106, generator=g)[:k]
        att[0, 2:6] = torch.tensor([5, -1, 9, 9])                       # entries after the first -1 are ignored
        samples.append((torch.randn(3, 64, 64, generator=g), torch.randint(1, 179, (n,), generator=g), boxes, att))
    got = data.collate_on_device(samples, 106, "cuda")
    want = DO.collate(samples, 106)
    names = ("imgs", "objs", "boxes", "masks", "obj_to_img", "attribute", "masks_shift", "boxes_shift")
    for n_, a, r in zip(names, got, want):
        assert a.dtype == r.dtype and a.shape == r.shape, n_
        assert torch.equal(a.cpu(), r), n_
    assert not got[4].is_cuda and got[3].is_cuda
    # the collated batch drives the generator exactly like a loader batch
    ts = TrainStep(64, device="cuda")
    out = ts.netG(got[0], got[1], got[2], got[3], got[4], torch.randn(got[1].shape[0], 64, device="cuda"), got[5], got[6],
                  got[7], got[5])
    assert out[4].shape == (6, 3, 64, 64) and torch.isfinite(out[4]).all()
