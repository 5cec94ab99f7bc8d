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


def oracle_step(model, batch, seeds=(123, 124), swap=None):
    """train64.py:141-370 through the oracle; seeds pin the CropEncoder noise of the two generator forwards.
    swap = (matrix, random.Random): apply the GT-attribute swap of train64.py:169-188 after the attribute estimation."""
    b = dict(batch)
    b["attribute_GT"] = b["attribute"].clone()
    nets = model.nets()
    with torch.no_grad():
        crops = O.crop_bbox_batch(b["imgs"], b["boxes"], b["obj_to_img"], model.obj_size)
    est = O.estimate_attributes(nets["att"](crops).detach(), b["attribute"])
    if swap is not None:
        b["attribute"], est, _ = O.swap_attributes(b["attribute"], est, b["objs"], b["obj_to_img"], b["imgs"].shape[0],
                                                   swap[0], swap[1])
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


_NETS = (("G", "netG", "G"), ("D_img", "netD_image", "D_img"), ("D_obj", "netD_object", "D_obj"), ("D_att", "netD_att", "D_att"))


class SyncedOracle:
    """The oracle + torch.optim.Adam (train64.py:111-114), re-synchronised to a TrainStep's parameters, buffers and Adam
    moments before every iteration ("teacher forcing").  The reference's training dynamics are chaotic at random init
    (Adam's first steps move every weight by +-lr whatever the gradient's size: the oracle run with 1 and with 8 host
    threads already differs by 3e-3 in the images after one update and 1.5e-1 after two), so free-running trajectories
    cannot be compared tightly; one iteration from IDENTICAL state can — and that is exactly the property a stale weight
    operand breaks (iteration k+1 must compute with the weights iteration k wrote)."""

    def __init__(self, size, states, lr=2e-4):
        self.model = O.OracleModel(size, 0, states)
        self.opts = {n: torch.optim.Adam([v for v in getattr(self.model, a).values() if v.requires_grad], lr=lr,
                                         betas=(0.5, 0.999)) for n, _, a in _NETS}

    def sync_from(self, ts):
        ours_opt = dict(G=ts.opt_G, D_img=ts.opt_D[0], D_obj=ts.opt_D[1], D_att=ts.opt_D[2])
        for n, attr, a in _NETS:
            net, st, opt = getattr(ts, attr), getattr(self.model, a), self.opts[n]
            with torch.no_grad():
                for k, v in net.state_dict().items():
                    st[k].copy_(v.detach().cpu())
            for k, p in net.named_parameters():
                so = ours_opt[n].state.get(p)
                if so:
                    opt.state[st[k]] = dict(step=torch.tensor(float(so["step"])), exp_avg=so["exp_avg"].detach().cpu().clone(),
                                            exp_avg_sq=so["exp_avg_sq"].detach().cpu().clone())

    def params(self):
        return {n: {k: v.detach().clone() for k, v in getattr(self.model, a).items() if v.requires_grad} for n, _, a in _NETS}

    def iterate(self, batch, seeds):
        """one iteration of train64.py:141-370 incl. the Adam updates; returns losses and the generator outputs"""
        model, opts = self.model, self.opts
        b = dict(batch)
        b.setdefault("attribute_GT", b["attribute"].clone())
        nets = model.nets()
        with torch.no_grad():
            crops = O.crop_bbox_batch(b["imgs"], b["boxes"], b["obj_to_img"], model.obj_size)
        est = O.estimate_attributes(nets["att"](crops).detach(), b["attribute"])
        torch.manual_seed(seeds[0])
        out_d = model.generator(b, est)
        d_loss, _ = O.d_step_loss(nets, b, out_d, model.pos_weight)
        model.zero_grad((model.D_img, model.D_obj, model.D_att))
        d_loss.backward()
        for n in ("D_img", "D_obj", "D_att"):
            opts[n].step()
        torch.manual_seed(seeds[1])
        out_g = model.generator(b, est)
        g_loss, _ = O.g_step_loss(nets, b, out_g, model.pos_weight)
        model.zero_grad((model.G,))
        g_loss.backward()
        opts["G"].step()
        return dict(d_loss=float(d_loss.detach()), g_loss=float(g_loss.detach()), out_g=[t.detach().clone() for t in out_g])


def reference_eps(seed, n_obj, z_dim=64):
    """the three CropEncoder noise draws of one generator forward (generator_obj_att.py:620, 640, 645) for a seed"""
    torch.manual_seed(seed)
    return [torch.randn(n_obj, z_dim) for _ in range(3)]


def run_synced_training(ts, batch, size, states, n_steps, img_tol, loss_tol, cos_min, seed0=5, verbose=False, step_fn=None):
    """n_steps iterations of ts.step(optimizer_step=True), each compared with ONE oracle iteration started from ts's own
    state before that iteration: generator outputs (rel-L2 <= img_tol), both losses (<= loss_tol) and, per network, the
    direction of the parameter update (cosine >= cos_min[name]; elements whose gradient sits at the fp32 noise floor
    legitimately move by -+lr).  Also asserts that every iteration's images differ from the previous one's (the update is
    visible).  step_fn(b, seeds): replaces ts.step (e.g. a CUDA-graph replay).  Returns the per-step measurements."""
    so = SyncedOracle(size, states)
    b = ts.to_device(batch)
    prev_img, log = None, []
    for it in range(n_steps):
        so.sync_from(ts)
        before = so.params()
        seeds = (seed0 + it, 10 * seed0 + it)
        r = step_fn(b, seeds) if step_fn is not None else ts.step(b, optimizer_step=True, seeds=seeds)
        ref = so.iterate(batch, seeds)
        m = {}
        for i in (0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10):
            m["out_g[%d]" % i] = rel(r["out_g"][i], ref["out_g"][i])
        m["d_loss"] = abs(float(r["d_loss"]) - ref["d_loss"]) / abs(ref["d_loss"])
        m["g_loss"] = abs(float(r["g_loss"]) - ref["g_loss"]) / abs(ref["g_loss"])
        for n, attr, a in _NETS:
            net, st = getattr(ts, attr), getattr(so.model, a)
            ua = torch.cat([(p.detach().cpu().double() - before[n][k].double()).reshape(-1) for k, p in net.named_parameters()])
            ur = torch.cat([(st[k].detach().double() - before[n][k].double()).reshape(-1) for k, _ in net.named_parameters()])
            assert float(ua.abs().max()) > 0 and float(ur.abs().max()) > 0, (n, "parameters did not move")
            m["upd_cos_" + n] = float(torch.nn.functional.cosine_similarity(ua, ur, dim=0))
        if prev_img is not None:
            m["change"] = rel(r["out_g"][4], prev_img)
        prev_img = r["out_g"][4].detach().clone()
        log.append(m)
        if verbose:
            print("iteration %d: " % it + " ".join("%s=%.3g" % kv for kv in m.items()))
    for it, m in enumerate(log):
        for k, v in m.items():
            if k.startswith("out_g"):
                assert v <= img_tol, (it, k, v)
            elif k.endswith("_loss"):
                assert v <= loss_tol, (it, k, v)
            elif k.startswith("upd_cos_"):
                assert v >= cos_min[k[8:]], (it, k, v)
            elif k == "change":
                assert v > 50 * img_tol or v > 2e-2, (it, "the optimizer update is not visible in the next iteration", v)
    return log


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
