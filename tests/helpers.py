"""Shared test helpers: run the oracle step with the golden RNG protocol, compare gradients, load states."""
import os
import zlib

import torch

from oracle import gan_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def proj(key, t):
    g = torch.Generator()
    g.manual_seed(zlib.crc32(key.encode()) & 0x7FFFFFFF)
    r = torch.randn(t.numel(), generator=g, dtype=torch.float64)
    return float((t.detach().double().cpu().view(-1) * r).sum() / max(1.0, t.numel() ** 0.5))


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def oracle_step(model, batch, seeds=(123, 124)):
    """train64.py:141-370 through the oracle; seeds pin the CropEncoder noise of the two generator forwards."""
    b = dict(batch)
    b["attribute_GT"] = b["attribute"].clone()
    nets = model.nets()
    with torch.no_grad():
        crops = O.crop_bbox_batch(b["imgs"], b["boxes"], b["obj_to_img"], model.obj_size)
    est = O.estimate_attributes(nets["att"](crops).detach(), b["attribute"])
    torch.manual_seed(seeds[0])
    out_d = model.generator(b, est)
    d_loss, d_terms = O.d_step_loss(nets, b, out_d, model.pos_weight)
    model.zero_grad((model.D_img, model.D_obj, model.D_att))
    d_loss.backward()
    d_grads = {n: {k: v.grad.clone() for k, v in st.items() if v.requires_grad} for n, st in
               (("D_img", model.D_img), ("D_obj", model.D_obj), ("D_att", model.D_att))}
    torch.manual_seed(seeds[1])
    out_g = model.generator(b, est)
    g_loss, g_terms = O.g_step_loss(nets, b, out_g, model.pos_weight)
    model.zero_grad((model.G,))
    g_loss.backward()
    g_grads = {k: v.grad.clone() for k, v in model.G.items() if v.requires_grad}
    return dict(d_loss=d_loss.detach(), g_loss=g_loss.detach(), d_terms=d_terms, g_terms=g_terms, d_grads=d_grads,
                g_grads=g_grads, out_d=[t.detach() for t in out_d], out_g=[t.detach() for t in out_g], attribute_est=est)


def load_states(ts, states):
    for net, key in ((ts.netG, "G"), (ts.netD_image, "D_img"), (ts.netD_object, "D_obj"), (ts.netD_att, "D_att")):
        net.load_state_dict(states[key], strict=True)


def check_step_against(ts, res, ref, img_tol, loss_tol, grad_tol, cos_min, zero_tol=1e-5):
    """ts: TrainStep after step(optimizer_step=False); ref: dict in oracle_step format."""
    errs = {}
    for i in range(11):
        errs["out_g[%d]" % i] = rel(res["out_g"][i], ref["out_g"][i])
        assert errs["out_g[%d]" % i] <= img_tol, ("generator output %d" % i, errs)
    assert abs(float(res["d_loss"]) - float(ref["d_loss"])) <= loss_tol * abs(float(ref["d_loss"]))
    assert abs(float(res["g_loss"]) - float(ref["g_loss"])) <= loss_tol * abs(float(ref["g_loss"]))
    ga, gr = [], []
    for k, p in ts.netG.named_parameters():
        r = ref["g_grads"][k]
        ga.append(p.grad.detach().cpu().reshape(-1))
        gr.append(r.reshape(-1))
        if float(r.norm()) < 1e-6:     # analytically zero gradients (Linear bias feeding BatchNorm1d)
            assert float(p.grad.abs().max()) <= zero_tol, k
        else:
            e = rel(p.grad, r)
            assert e <= grad_tol, ("G grad", k, e)
    ga, gr = torch.cat(ga).double(), torch.cat(gr).double()
    cos = float(torch.nn.functional.cosine_similarity(ga, gr, dim=0))
    assert cos >= cos_min, ("G grad cosine", cos)
    for name, net in (("D_img", ts.netD_image), ("D_obj", ts.netD_object), ("D_att", ts.netD_att)):
        for k, p in net.named_parameters():
            e = rel(p.grad, ref["d_grads"][name][k])
            assert e <= grad_tol, (name, k, e)
    return cos
