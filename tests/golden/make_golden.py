"""Generate golden vectors from the UNMODIFIED reference (build container only: needs /root/reference).

For each image size the script
  1. builds the oracle's seeded states and loads them (strict) into the reference's own modules,
  2. runs one D+G step (wiring of train64.py:141-370, losses shared with the oracle) through the
     reference classes on CPU,
  3. runs the same step through oracle/gan_oracle.py and reports the agreement,
  4. stores small fixtures: outputs, logits, losses, updated BN/SN buffers and, per parameter, the
     gradient's L2 norm plus a seeded random projection (so 30 M-element gradients pin to a few floats).

Run:  python tests/golden/make_golden.py            (writes tests/golden/step64.pt, step128.pt, crop.pt)
"""
import os
import sys
import zlib

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from oracle import gan_oracle as O  # noqa: E402


def proj(key, t):
    g = torch.Generator()
    g.manual_seed(zlib.crc32(key.encode()) & 0x7FFFFFFF)
    r = torch.randn(t.numel(), generator=g, dtype=torch.float64)
    return float((t.detach().double().view(-1) * r).sum() / max(1.0, t.numel() ** 0.5))


def grad_summary(grads):
    return {k: (float(v.double().norm()), proj(k, v)) for k, v in grads.items()}


def build_reference(image_size, states):
    if image_size == 64:
        from models.generator_obj_att import Generator
        from models.discriminator import AttributeDiscriminator as AttD
    else:
        from models.generator_obj_att128 import Generator
        from models.discriminator import AttributeDiscriminator128 as AttD
    from models.discriminator import ImageDiscriminator, ObjectDiscriminator, add_sn
    G = Generator(num_embeddings=O.NUM_OBJECTS, obj_att_dim=64, z_dim=64, clstm_layers=3, obj_size=image_size // 2,
                  attribute_dim=O.NUM_ATTRIBUTES)
    D_img = add_sn(ImageDiscriminator(conv_dim=64))
    D_obj = add_sn(ObjectDiscriminator(n_class=O.NUM_OBJECTS))
    D_att = add_sn(AttD(n_attribute=O.NUM_ATTRIBUTES))
    for net, key in ((G, "G"), (D_img, "D_img"), (D_obj, "D_obj"), (D_att, "D_att")):
        ref_sd = net.state_dict()
        assert set(ref_sd.keys()) == set(states[key].keys()), (key, set(ref_sd.keys()) ^ set(states[key].keys()))
        for k in ref_sd:
            assert tuple(ref_sd[k].shape) == tuple(states[key][k].shape), (key, k, ref_sd[k].shape, states[key][k].shape)
        net.load_state_dict({k: v.clone() for k, v in states[key].items()}, strict=True)
        net.train()
    return G, D_img, D_obj, D_att


def reference_step(nets4, batch, image_size):
    G, D_img, D_obj, D_att = nets4
    from models.bilinear import crop_bbox_batch
    b = dict(batch)
    b["attribute_GT"] = b["attribute"].clone()
    nets = dict(image=D_img, object=lambda x: D_obj(x, b["objs"]), att=D_att)
    pw = O.pos_weight_vector()
    crops = crop_bbox_batch(b["imgs"], b["boxes"], b["obj_to_img"], image_size // 2)
    est = O.estimate_attributes(D_att(crops).detach(), b["attribute"])

    def gen():
        return G(b["imgs"], b["objs"], b["boxes"], b["masks"], b["obj_to_img"], b["z"], b["attribute"], b["masks_shift"],
                 b["boxes_shift"], est)

    torch.manual_seed(123)
    out_d = gen()
    d_loss, d_terms = O.d_step_loss(nets, b, out_d, pw)
    for n in (D_img, D_obj, D_att):
        n.zero_grad()
    d_loss.backward()
    d_grads = {name: {k: p.grad.clone() for k, p in n.named_parameters()} for name, n in
               (("D_img", D_img), ("D_obj", D_obj), ("D_att", D_att))}
    torch.manual_seed(124)
    out_g = gen()
    g_loss, g_terms = O.g_step_loss(nets, b, out_g, pw)
    G.zero_grad()
    g_loss.backward()
    g_grads = {k: p.grad.clone() for k, p in G.named_parameters()}
    return dict(d_loss=d_loss.detach(), g_loss=g_loss.detach(), d_terms=d_terms, g_terms=g_terms, d_grads=d_grads,
                g_grads=g_grads, out_d=[t.detach() for t in out_d], out_g=[t.detach() for t in out_g], attribute_est=est)


def oracle_step(model, batch):
    # same RNG protocol as reference_step: seed 123 before the D-step G forward, 124 before the G-step one
    b = dict(batch)
    b["attribute_GT"] = b["attribute"].clone()
    nets = model.nets()
    with torch.no_grad():
        crops = O.crop_bbox_batch(b["imgs"], b["boxes"], b["obj_to_img"], model.obj_size)
    est = O.estimate_attributes(nets["att"](crops).detach(), b["attribute"])
    torch.manual_seed(123)
    out_d = model.generator(b, est)
    d_loss, d_terms = O.d_step_loss(nets, b, out_d, model.pos_weight)
    model.zero_grad((model.D_img, model.D_obj, model.D_att))
    d_loss.backward()
    d_grads = {n: {k: v.grad.clone() for k, v in st.items() if v.requires_grad} for n, st in
               (("D_img", model.D_img), ("D_obj", model.D_obj), ("D_att", model.D_att))}
    torch.manual_seed(124)
    out_g = model.generator(b, est)
    g_loss, g_terms = O.g_step_loss(nets, b, out_g, model.pos_weight)
    model.zero_grad((model.G,))
    g_loss.backward()
    g_grads = {k: v.grad.clone() for k, v in model.G.items() if v.requires_grad}
    return dict(d_loss=d_loss.detach(), g_loss=g_loss.detach(), d_terms=d_terms, g_terms=g_terms, d_grads=d_grads,
                g_grads=g_grads, out_d=[t.detach() for t in out_d], out_g=[t.detach() for t in out_g], attribute_est=est)


def relerr(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def make_step(image_size, n_images, seed):
    torch.set_num_threads(8)
    states = O.make_states(image_size, seed)
    batch = O.synth_batch(n_images, image_size, objs_per_image=None, seed=seed + 7)
    ref_nets = build_reference(image_size, states)
    ref = reference_step(ref_nets, batch, image_size)
    model = O.OracleModel(image_size, seed, states)
    orc = oracle_step(model, batch)
    print("== %d^2  N=%d O=%d" % (image_size, n_images, batch["objs"].numel()))
    print("d_loss ref %.6f oracle %.6f | g_loss ref %.6f oracle %.6f" % (ref["d_loss"], orc["d_loss"], ref["g_loss"], orc["g_loss"]))
    for i, (a, b) in enumerate(zip(orc["out_g"], ref["out_g"])):
        print("  out_g[%d] rel err %.2e" % (i, relerr(a, b)))
    worst = 0.0
    for k, v in ref["g_grads"].items():
        e = relerr(orc["g_grads"][k], v) if float(v.norm()) > 1e-8 else float((orc["g_grads"][k] - v).abs().max())
        worst = max(worst, e)
    print("  worst G-grad per-tensor rel err %.2e" % worst)
    for n in ("D_img", "D_obj", "D_att"):
        w = max(relerr(orc["d_grads"][n][k], v) for k, v in ref["d_grads"][n].items())
        print("  worst %s-grad rel err %.2e" % (n, w))
    # buffers after the step (BN running stats, SN u/v) from the reference modules
    bufs = {}
    for name, net in zip(("G", "D_img", "D_obj", "D_att"), ref_nets):
        sd = net.state_dict()
        bufs[name] = {k: sd[k].clone() for k in sd if not O.is_parameter(k)}
        st = getattr(model, name)
        wb = max((relerr(st[k].float(), v.float()) for k, v in bufs[name].items()), default=0.0)
        print("  worst %s buffer rel err %.2e" % (name, wb))
    gold = dict(image_size=image_size, n_images=n_images, seed=seed, batch_seed=seed + 7,
                d_loss=float(ref["d_loss"]), g_loss=float(ref["g_loss"]),
                d_terms={k: float(v) for k, v in ref["d_terms"].items()},
                g_terms={k: float(v) for k, v in ref["g_terms"].items()},
                out_g=[t.clone() for t in ref["out_g"]],
                out_d_imgs=[t.clone() for t in ref["out_d"][4:7]],
                g_grads=grad_summary(ref["g_grads"]),
                d_grads={n: grad_summary(g) for n, g in ref["d_grads"].items()},
                buffers={n: {k: (v.clone() if v.numel() <= 2048 else (float(v.double().norm()), proj(k, v)))
                             for k, v in b.items()} for n, b in bufs.items()})
    return gold


def make_crop():
    from models.bilinear import crop_bbox_batch
    g = torch.Generator()
    g.manual_seed(5)
    feats = torch.randn(3, 3, 64, 64, generator=g)
    boxes = torch.tensor([[0.0, 0.0, 1.0, 1.0], [0.1, 0.2, 0.55, 0.9], [0.5, 0.5, 0.5, 0.5], [0.33, 0.0, 1.0, 0.41],
                          [0.0, 0.7, 0.3, 1.0], [0.25, 0.25, 0.75, 0.75], [0.9, 0.05, 1.0, 0.15]])
    b2f = torch.tensor([0, 0, 1, 1, 1, 2, 2])
    feats.requires_grad_(True)
    crops = crop_bbox_batch(feats, boxes, b2f, 32)
    w = torch.randn(crops.shape, generator=g)
    (crops * w).sum().backward()
    mine = O.crop_bbox_batch(feats.detach(), boxes, b2f, 32)
    print("crop oracle vs reference max abs diff %.2e" % float((mine - crops.detach()).abs().max()))
    # unsorted mapping exercises the inverse-permutation branch bilinear.py:99-104
    b2f_u = torch.tensor([2, 0, 1, 0, 2, 1, 0])
    crops_u = crop_bbox_batch(feats.detach(), boxes, b2f_u, 16, 24)
    return dict(feats=feats.detach().clone(), boxes=boxes, b2f=b2f, crops=crops.detach().clone(), w=w,
                dfeats=feats.grad.clone(), b2f_u=b2f_u, crops_u=crops_u.clone())


def make_swap():
    """The "change GT attribute" block of the training loop (train64.py:169-188) is inline code, not a function: its
    lines are read from the reference file and executed UNMODIFIED over a prepared namespace (synthetic co-occurrence
    matrix with the real one's shape/dtype, Python `random` seeded), and inputs + results are stored."""
    import math
    import random
    import textwrap
    lines = open("/root/reference/train64.py").read().splitlines()
    lo = next(i for i, l in enumerate(lines) if "# change GT attribute:" in l)
    hi = next(i for i, l in enumerate(lines) if "# Generate fake image" in l)
    code = textwrap.dedent("\n".join(lines[lo:hi]))
    cases = []
    for seed, n_images in ((0, 6), (1, 7), (2, 3), (3, 2)):
        batch = O.synth_batch(n_images, 64, objs_per_image=None, seed=40 + seed, sparse_attributes=True)
        g = torch.Generator().manual_seed(seed)
        matrix = torch.randint(0, 5000, (O.NUM_OBJECTS, O.NUM_ATTRIBUTES), generator=g).float()
        matrix[0] = 0
        attribute = batch["attribute"].clone()
        attribute_GT = attribute.clone()
        est_logits = torch.randn(attribute.shape, generator=g)
        attribute_est = O.estimate_attributes(est_logits, attribute)
        ns = dict(torch=torch, math=math, random=random, matrix=matrix, device=torch.device("cpu"), imgs=batch["imgs"],
                  objs=batch["objs"], obj_to_img=batch["obj_to_img"], attribute=attribute.clone(),
                  attribute_GT=attribute_GT, attribute_est=attribute_est.clone())
        random.seed(1000 + seed)
        exec(code, ns)
        cases.append(dict(seed=1000 + seed, n_images=n_images, objs=batch["objs"], obj_to_img=batch["obj_to_img"],
                          matrix=matrix, attribute_in=attribute, attribute_est_in=attribute_est,
                          attribute_out=ns["attribute"], attribute_est_out=ns["attribute_est"]))
        changed = int((ns["attribute"] != attribute).any(1).sum())
        print("swap case %d: N=%d O=%d rows changed %d" % (seed, n_images, attribute.shape[0], changed))
    return cases


def make_data():
    """data/utils.py:47-66 imagenet_deprocess_batch (imported from the reference, unmodified) on seeded images, with and
    without the per-image rescale; plus the loader's one-hot attribute construction (vg_custom_mask.py:160-171 restated by
    executing its own statements is not possible without h5py — the lines are quoted in the oracle instead)."""
    from data.utils import imagenet_deprocess_batch
    g = torch.Generator().manual_seed(11)
    imgs = torch.randn(5, 3, 16, 12, generator=g) * 1.3
    imgs[1] = imgs[1] * 0.01 + 0.7          # low contrast image: the rescale stretches it
    imgs[2, :, 0, 0] = 40.0                 # outlier: everything else lands near zero
    return dict(imgs=imgs, out_rescale=imagenet_deprocess_batch(imgs, rescale=True),
                out_plain=imagenet_deprocess_batch(imgs, rescale=False))


if __name__ == "__main__":
    out = os.path.dirname(os.path.abspath(__file__))
    torch.save(make_swap(), os.path.join(out, "swap.pt"))
    torch.save(make_data(), os.path.join(out, "data.pt"))
    if "--swap-only" in sys.argv:
        sys.exit(0)
    torch.save(make_crop(), os.path.join(out, "crop.pt"))
    torch.save(make_step(64, 2, 0), os.path.join(out, "step64.pt"))
    torch.save(make_step(128, 2, 0), os.path.join(out, "step128.pt"))
    print("golden vectors written")
