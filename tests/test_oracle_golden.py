"""not gpu: the oracle restatement against golden vectors produced by the UNMODIFIED reference
(tests/golden/make_golden.py, run in the build container where /root/reference exists)."""
import os

import pytest
import torch

from helpers import GOLD, oracle_step, proj, rel
from oracle import gan_oracle as O


def test_crop_oracle_matches_reference_golden():
    g = torch.load(os.path.join(GOLD, "crop.pt"))
    feats = g["feats"].clone().requires_grad_(True)
    crops = O.crop_bbox_batch(feats, g["boxes"], g["b2f"], 32)
    assert float((crops - g["crops"]).abs().max()) < 1e-6
    (crops * g["w"]).sum().backward()
    assert float((feats.grad - g["dfeats"]).abs().max()) < 1e-5
    cu = O.crop_bbox_batch(g["feats"], g["boxes"], g["b2f_u"], 16, 24)
    assert float((cu - g["crops_u"]).abs().max()) < 1e-6


def test_crop_taps_agree_with_grid_sample():
    """the explicit integer/weight restatement (crop_taps) reproduces grid_sample's output"""
    g = torch.load(os.path.join(GOLD, "crop.pt"))
    feats, boxes, b2f = g["feats"], g["boxes"], g["b2f"]
    ix0, iy0, fx, fy = O.crop_taps(boxes, 64, 64, 32, 32)
    B = boxes.shape[0]
    out = torch.zeros(B, 3, 32, 32)
    for b in range(B):
        img = feats[b2f[b]]
        for i in range(32):
            for j in range(32):
                acc = torch.zeros(3)
                for (yy, wy) in ((int(iy0[b, i]), 1 - fy[b, i]), (int(iy0[b, i]) + 1, fy[b, i])):
                    for (xx, wx) in ((int(ix0[b, j]), 1 - fx[b, j]), (int(ix0[b, j]) + 1, fx[b, j])):
                        if 0 <= yy < 64 and 0 <= xx < 64:
                            acc = acc + img[:, yy, xx] * (wy * wx)
                out[b, :, i, j] = acc
    assert float((out - g["crops"]).abs().max()) < 2e-6


def test_rasterise_and_shift_contract():
    boxes = torch.tensor([[0.1, 0.2, 0.3, 0.9], [0.0, 0.0, 1.0, 1.0], [0.7, 0.1, 0.95, 0.5], [0.2578125, 0.0, 0.5078125, 1.0]])
    m = O.rasterize_boxes(boxes, 64, 64)
    assert m.shape == (4, 1, 64, 64) and float(m[1].sum()) == 64 * 64
    # python round is half-to-even on double: 0.2578125*64 = 16.5 -> 16, 0.5078125*64 = 32.5 -> 32
    assert float(m[3, 0, 0].sum()) == 16 and float(m[3, 0, 0, 16]) == 1 and float(m[3, 0, 0, 15]) == 0
    s = O.shift_boxes(boxes)
    assert torch.allclose(s[0], torch.tensor([0.1 + 0.7 * 0.8, 0.2, 0.3 + 0.7 * 0.8, 0.9]))
    assert torch.equal(s[1], boxes[1])
    assert torch.allclose(s[2], torch.tensor([0.7 - 0.7 * 0.8, 0.1, 0.95 - 0.7 * 0.8, 0.5]))


@pytest.mark.parametrize("size", [64, 128])
def test_oracle_step_matches_reference_golden(size):
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    gold = torch.load(os.path.join(GOLD, "step%d.pt" % size))
    states = O.make_states(size, gold["seed"])
    batch = O.synth_batch(gold["n_images"], size, objs_per_image=None, seed=gold["batch_seed"])
    model = O.OracleModel(size, gold["seed"], states)
    res = oracle_step(model, batch)
    assert abs(float(res["d_loss"]) - gold["d_loss"]) < 1e-4 * abs(gold["d_loss"])
    assert abs(float(res["g_loss"]) - gold["g_loss"]) < 1e-4 * abs(gold["g_loss"])
    for a, r in zip(res["out_g"], gold["out_g"]):
        assert rel(a, r) < 2e-5
    for a, r in zip(res["out_d"][4:7], gold["out_d_imgs"]):
        assert rel(a, r) < 2e-5
    # gradients: norm + seeded projection per parameter (noise floor of end-to-end G grads is ~5e-3)
    for k, (nrm, pr) in gold["g_grads"].items():
        gk = res["g_grads"][k]
        if nrm < 1e-6:
            assert float(gk.norm()) < 1e-5
            continue
        assert abs(float(gk.double().norm()) - nrm) <= 2e-2 * nrm, k
        assert abs(proj(k, gk) - pr) <= 2e-2 * nrm / max(1.0, gk.numel() ** 0.5) * 8 + 1e-7, k
    for n, gs in gold["d_grads"].items():
        for k, (nrm, pr) in gs.items():
            gk = res["d_grads"][n][k]
            assert abs(float(gk.double().norm()) - nrm) <= 1e-3 * nrm + 1e-9, (n, k)
    for n, bufs in gold["buffers"].items():
        st = getattr(model, n)
        for k, v in bufs.items():
            if isinstance(v, tuple):
                assert abs(float(st[k].double().norm()) - v[0]) <= 1e-4 * v[0] + 1e-9, (n, k)
            else:
                assert rel(st[k].float(), v.float()) < 1e-4 or float((st[k].float() - v.float()).abs().max()) < 1e-6, (n, k)


def test_attribute_swap_matches_reference_lines():
    """train64.py:169-188 executed unmodified (make_golden.make_swap) vs the oracle restatement, incl. the Python RNG
    stream (randrange before choices), images with < 2 objects, un-annotated objects and N < 3"""
    import random
    cases = torch.load(os.path.join(GOLD, "swap.pt"))
    assert len(cases) == 4
    for c in cases:
        rng = random.Random(c["seed"])
        att, est, rows = O.swap_attributes(c["attribute_in"], c["attribute_est_in"], c["objs"], c["obj_to_img"],
                                           c["n_images"], c["matrix"], rng)
        assert torch.equal(att, c["attribute_out"]) and torch.equal(est, c["attribute_est_out"])
        changed = (c["attribute_out"] != c["attribute_in"]).any(1).nonzero().view(-1)
        assert set(changed.tolist()) <= set(rows.tolist())


def test_masks_to_layout_restatement_is_self_consistent():
    """oracle/layout_oracle.py (sg2im semantics; the reference ships only the call site, utils/draw_box.py:482-483): the
    explicit integer/weight restatement (layout_taps) reproduces the grid_sample formulation, objects of different images do
    not mix, and the result is the per-image sum of single-object layouts"""
    from oracle import layout_oracle as LO
    g = torch.Generator().manual_seed(0)
    O_, D, M, H, W = 5, 3, 4, 9, 7
    vecs, masks = torch.randn(O_, D, generator=g), torch.rand(O_, M, M, generator=g)
    xy0 = torch.rand(O_, 2, generator=g) * 0.5
    boxes = torch.cat([xy0, xy0 + 0.2 + 0.3 * torch.rand(O_, 2, generator=g)], 1)
    o2i = torch.tensor([0, 0, 2, 2, 2])
    out = LO.masks_to_layout(vecs, boxes, masks, o2i, H, W, N=3)
    assert out.shape == (3, D, H, W) and float(out[1].abs().max()) == 0.0
    ix0, iy0, fx, fy = LO.layout_taps(boxes, M, H, W)
    want = torch.zeros(3, D, H, W)
    for o in range(O_):
        for y in range(H):
            for x in range(W):
                s = 0.0
                for (yy, wy) in ((int(iy0[o, y]), 1 - fy[o, y]), (int(iy0[o, y]) + 1, fy[o, y])):
                    for (xx, wx) in ((int(ix0[o, x]), 1 - fx[o, x]), (int(ix0[o, x]) + 1, fx[o, x])):
                        if 0 <= yy < M and 0 <= xx < M:
                            s = s + float(masks[o, yy, xx]) * float(wy * wx)
                want[o2i[o], :, y, x] += vecs[o] * s
    assert float((out - want).abs().max()) < 1e-5
    single = sum(LO.masks_to_layout(vecs[o:o + 1], boxes[o:o + 1], masks[o:o + 1], torch.zeros(1, dtype=torch.long), H, W, N=1)
                 for o in (2, 3, 4))
    assert float((single[0] - out[2]).abs().max()) < 1e-6


def test_deprocess_restatement_matches_reference_golden():
    """data/utils.py:47-66 (unmodified reference function, make_golden.make_data) vs oracle/data_oracle.py, bit for bit"""
    from oracle import data_oracle as DO
    g = torch.load(os.path.join(GOLD, "data.pt"))
    assert torch.equal(DO.imagenet_deprocess_batch(g["imgs"], True), g["out_rescale"])
    assert torch.equal(DO.imagenet_deprocess_batch(g["imgs"], False), g["out_plain"])
    att = torch.tensor([[3, 5, -1, 7], [-1, 2, 2, 2], [0, 105, 4, 9]])
    oh = DO.one_hot_attributes(att, 106)
    assert oh[0].nonzero().view(-1).tolist() == [3, 5] and float(oh[1].sum()) == 0 and oh[2].nonzero().view(-1).tolist() == [0, 4, 9, 105]
