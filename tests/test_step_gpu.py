"""-m gpu: the full D+G step through the CUDA kernels against the CPU oracle (same seeded inputs) and against the
golden vectors the unmodified reference produced.  Stated tolerances (SURVEY.md App. D noise floors):
  fp32 mode : images / crops / z  rel-L2 <= 1e-4, losses 1e-4, per-parameter gradient rel-L2 <= 2e-2 with global
              cosine >= 0.9999 (fp32 end-to-end G-gradients are chaotic at the 5e-3 level even reference-vs-reference)
  bf16 mode : images rel-L2 <= 6e-2, losses 5e-2, G-gradient cosine >= 0.85 (tcgen05 bf16 operands, fp32 accumulate;
              measured 0.906 on B200 — the reference's own CPU bf16 autocast reaches 0.970)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

from helpers import GOLD, check_step_against, load_states, oracle_step, rel  # noqa: E402
from b200gan import ops  # noqa: E402
from b200gan.step import TrainStep  # noqa: E402
from oracle import gan_oracle as O  # noqa: E402


def _run(size, precision, n_images=2, batch_seed=7, objs_per_image=None):
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    states = O.make_states(size, 0)
    batch = O.synth_batch(n_images, size, objs_per_image, batch_seed)
    ops.set_precision(precision)
    try:
        ts = TrainStep(size, device="cuda")
        load_states(ts, states)
        res = ts.step(ts.to_device(batch), optimizer_step=False, seeds=(123, 124))
        torch.cuda.synchronize()
    finally:
        ops.set_precision("fp32")
    model = O.OracleModel(size, 0, states)
    ref = oracle_step(model, batch)
    return ts, res, model, ref


@pytest.mark.parametrize("size", [64, 128])
def test_fp32_step_matches_oracle_and_reference_golden(size):
    ts, res, model, ref = _run(size, "fp32")
    cos = check_step_against(ts, res, ref, img_tol=1e-4, loss_tol=1e-4, grad_tol=2e-2, cos_min=0.9999)
    print("fp32 %d: G-grad cosine %.7f" % (size, cos))
    gold = torch.load(os.path.join(GOLD, "step%d.pt" % size))
    for a, r in zip(res["out_g"], gold["out_g"]):
        assert rel(a, r) < 1e-4
    assert abs(float(res["g_loss"]) - gold["g_loss"]) < 1e-4 * abs(gold["g_loss"])
    assert abs(float(res["d_loss"]) - gold["d_loss"]) < 1e-4 * abs(gold["d_loss"])
    # BN running statistics / SN power-iteration state after the step
    for name, net, st in (("G", ts.netG, model.G), ("D_img", ts.netD_image, model.D_img), ("D_obj", ts.netD_object, model.D_obj),
                          ("D_att", ts.netD_att, model.D_att)):
        for k, v in net.state_dict().items():
            if not O.is_parameter(k):
                assert rel(v.float(), st[k].float()) < 1e-4 or float((v.float().cpu() - st[k].float()).abs().max()) < 1e-5, (name, k)


def _module_cosines(named_grads, ref_grads, depth=2):
    groups = {}
    for k, g in named_grads:
        groups.setdefault(".".join(k.split(".")[:depth]), []).append((g.detach().cpu().reshape(-1).double(), ref_grads[k].reshape(-1).double()))
    out = {}
    for m, pairs in groups.items():
        a, r = torch.cat([p[0] for p in pairs]), torch.cat([p[1] for p in pairs])
        out[m] = float(torch.nn.functional.cosine_similarity(a, r, dim=0))
    return out


@pytest.mark.parametrize("size", [64, 128])
def test_bf16_step_within_stated_bound(size):
    """bf16 mode (tcgen05 operands + activation storage in bf16, everything else fp32) against the fp32 oracle, 64x64 and
    128x128.  Stated bounds: images rel-L2 <= 6e-2, losses 5e-2; discriminator gradients (D-step): per-network cosine >= 0.995
    and every weight tensor's cosine >= 0.97; generator gradients (G-step, through three bf16 discriminators): global
    cosine >= 0.85 (measured 0.881 / 0.869) and every sub-module's cosine >= 0.75 (measured: the weakest is
    attribute_encoder.bn0 — 256 parameters fed by 14 objects — at 0.854 / 0.811).  Reference level: the reference algorithm under torch's own bf16
    autocast reaches 0.867 / 3.6e-2 on this step (tests/test_wiring_cpu.py::test_step_wiring_bf16_operand_routing)."""
    ts, res, model, ref = _run(size, "bf16")
    for i in (4, 5, 6):
        assert rel(res["out_g"][i], ref["out_g"][i]) < 6e-2
    for i in (7, 8, 9, 10):
        assert rel(res["out_g"][i], ref["out_g"][i]) < 8e-2
    assert abs(float(res["d_loss"]) - float(ref["d_loss"])) < 5e-2 * abs(float(ref["d_loss"]))
    assert abs(float(res["g_loss"]) - float(ref["g_loss"])) < 5e-2 * abs(float(ref["g_loss"]))
    cosf = torch.nn.functional.cosine_similarity
    for name, net in (("D_img", ts.netD_image), ("D_obj", ts.netD_object), ("D_att", ts.netD_att)):
        a = torch.cat([p.grad.reshape(-1).cpu() for _, p in net.named_parameters()]).double()
        r = torch.cat([ref["d_grads"][name][k].reshape(-1) for k, _ in net.named_parameters()]).double()
        c = float(cosf(a, r, dim=0))
        print("bf16 %d: %s D-step gradient cosine %.5f" % (size, name, c))
        assert c > 0.995, (name, c)
        for k, p in net.named_parameters():
            r = ref["d_grads"][name][k]
            if k.endswith("weight_orig") and float(r.norm()) > 1e-6:
                ck = float(cosf(p.grad.reshape(-1).cpu().double(), r.reshape(-1).double(), dim=0))
                assert ck > 0.97, (name, k, ck)
    ga = torch.cat([p.grad.reshape(-1).cpu() for _, p in ts.netG.named_parameters()]).double()
    gr = torch.cat([ref["g_grads"][k].reshape(-1) for k, _ in ts.netG.named_parameters()]).double()
    cos = float(cosf(ga, gr, dim=0))
    per = _module_cosines([(k, p.grad) for k, p in ts.netG.named_parameters()], ref["g_grads"])
    worst = min(per.items(), key=lambda kv: kv[1])
    print("bf16 %d: G-grad cosine %.5f; worst sub-module %s %.4f" % (size, cos, worst[0], worst[1]))
    assert cos > 0.85
    assert worst[1] > 0.75, worst


def test_tf32_step_within_stated_bound():
    """tf32 mode: fp32 tensors everywhere, convolutions / linears with >= 32-aligned channels on tcgen05 kind::tf32 (the
    "fp32" half of BASELINE config 2 on the tensor cores).  Stated bounds vs the fp32 oracle: images / crops / z rel-L2 <=
    5e-3, losses 2e-3, discriminator gradient cosine >= 0.9999, generator gradient cosine >= 0.98 (2 images; 0.988 measured on
    the CPU emulation of the kernels, against 0.89 for bf16 on the same step)."""
    ts, res, model, ref = _run(64, "tf32")
    errs = [rel(res["out_g"][i], ref["out_g"][i]) for i in range(11)]
    print("tf32 64: output rel-L2 max %.2e" % max(errs))
    assert max(errs) < 5e-3, errs
    assert abs(float(res["d_loss"]) - float(ref["d_loss"])) < 2e-3 * abs(float(ref["d_loss"]))
    assert abs(float(res["g_loss"]) - float(ref["g_loss"])) < 2e-3 * abs(float(ref["g_loss"]))
    cosf = torch.nn.functional.cosine_similarity
    for name, net in (("D_img", ts.netD_image), ("D_obj", ts.netD_object), ("D_att", ts.netD_att)):
        a = torch.cat([p.grad.reshape(-1).cpu() for _, p in net.named_parameters()]).double()
        r = torch.cat([ref["d_grads"][name][k].reshape(-1) for k, _ in net.named_parameters()]).double()
        c = float(cosf(a, r, dim=0))
        print("tf32 64: %s D-step gradient cosine %.6f" % (name, c))
        assert c > 0.9999, (name, c)
    ga = torch.cat([p.grad.reshape(-1).cpu() for _, p in ts.netG.named_parameters()]).double()
    gr = torch.cat([ref["g_grads"][k].reshape(-1) for k, _ in ts.netG.named_parameters()]).double()
    c = float(cosf(ga, gr, dim=0))
    print("tf32 64: G-grad cosine %.6f" % c)
    assert c > 0.98


@pytest.mark.parametrize("precision", ["fp32", "tf32", "bf16"])
def test_config2_full_size_against_oracle(precision):
    """BASELINE config 2 at FULL size (64x64, batch 32, 8 objects/image = 256 objects) against the CPU oracle (~25 s on the
    host cores): outputs, losses, all parameter gradients.  fp32: images 1e-4, losses 1e-4, per-network gradient cosine >=
    0.9999 (D) / 0.999 (G); bf16: the bounds of test_bf16_step_within_stated_bound."""
    ts, res, model, ref = _run(64, precision, n_images=32, batch_seed=3, objs_per_image=8)
    fp32 = precision == "fp32"
    tol_out = {"fp32": 1e-4, "tf32": 5e-3, "bf16": 8e-2}[precision]
    tol_loss = {"fp32": 1e-4, "tf32": 2e-3, "bf16": 5e-2}[precision]
    cos_d = {"fp32": 0.9999, "tf32": 0.9999, "bf16": 0.995}[precision]
    cos_g = {"fp32": 0.999, "tf32": 0.995, "bf16": 0.85}[precision]
    for i in range(11):
        assert rel(res["out_g"][i], ref["out_g"][i]) < tol_out, i
    for k in ("d_loss", "g_loss"):
        assert abs(float(res[k]) - float(ref[k])) < tol_loss * abs(float(ref[k])), k
    cosf = torch.nn.functional.cosine_similarity
    for name, net in (("D_img", ts.netD_image), ("D_obj", ts.netD_object), ("D_att", ts.netD_att)):
        a = torch.cat([p.grad.reshape(-1).cpu() for _, p in net.named_parameters()]).double()
        r = torch.cat([ref["d_grads"][name][k].reshape(-1) for k, _ in net.named_parameters()]).double()
        c = float(cosf(a, r, dim=0))
        print("config 2 %s: %s gradient cosine %.6f" % (precision, name, c))
        assert c > cos_d, (name, c)
    ga = torch.cat([p.grad.reshape(-1).cpu() for _, p in ts.netG.named_parameters()]).double()
    gr = torch.cat([ref["g_grads"][k].reshape(-1) for k, _ in ts.netG.named_parameters()]).double()
    c = float(cosf(ga, gr, dim=0))
    print("config 2 %s: G gradient cosine %.6f" % (precision, c))
    assert c > cos_g


@pytest.mark.parametrize("optimizer", ["b200", "torch_fused", "torch"])
def test_three_training_iterations_follow_oracle(optimizer):
    """train64.py:254-262, 366-370: the weights move every iteration and every GEMM must see them.  Three full iterations
    (D-step, 3 x Adam, G-step on the updated discriminators, Adam) in fp32; each is compared with one oracle +
    torch.optim.Adam iteration started from the same state (helpers.SyncedOracle explains why not free-running).  The raw-
    pointer kernel optimizer and torch's fused Adam do not move tensor versions — round 1's packed operands went stale
    exactly there."""
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    from helpers import run_synced_training
    states = O.make_states(64, 0)
    batch = O.synth_batch(2, 64, 4, 7)
    ops.set_precision("fp32")
    ts = TrainStep(64, device="cuda", optimizer=optimizer)
    load_states(ts, states)
    log = run_synced_training(ts, batch, 64, states, 3, img_tol=1e-4, loss_tol=1e-4,
                              cos_min=dict(G=0.97, D_img=0.999, D_obj=0.999, D_att=0.999), verbose=True)
    assert len(log) == 3


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_cuda_graph_replays_train(precision):
    """The bench path: ONE captured CUDA graph of the whole iteration (incl. the Adam updates and the re-pack of every GEMM
    operand) replayed three times must train like the eager step: every replay is compared with one oracle iteration from
    the state the replay started in.  The CropEncoder noise comes from static device buffers filled with the reference's
    CPU-RNG draws before each replay."""
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    from helpers import reference_eps, run_synced_training
    states = O.make_states(64, 0)
    batch = O.synth_batch(2, 64, 4, 7)
    n_obj = batch["objs"].shape[0]
    ops.set_precision(precision)
    try:
        ts = TrainStep(64, device="cuda", optimizer="b200")
        load_states(ts, states)
        # per iteration the crop encoder is called (O) then (2O) in each of the two generator forwards
        bufs = [torch.zeros(n_obj, 64, device="cuda"), torch.zeros(2 * n_obj, 64, device="cuda"),
                torch.zeros(n_obj, 64, device="cuda"), torch.zeros(2 * n_obj, 64, device="cuda")]
        calls = [0]

        def eps_source(o, z, dev):
            t = bufs[calls[0] % 4]
            calls[0] += 1
            assert t.shape == (o, z)
            return t

        def fill(seeds):
            for half, seed in enumerate(seeds):
                e = reference_eps(seed, n_obj)
                bufs[2 * half].copy_(e[0])
                bufs[2 * half + 1].copy_(torch.cat(e[1:]))

        ts.netG.crop_encoder.eps_source = eps_source
        b = ts.to_device(batch)
        sd0 = {n: {k: v.clone() for k, v in net.state_dict().items()} for n, net in
               (("G", ts.netG), ("D_img", ts.netD_image), ("D_obj", ts.netD_object), ("D_att", ts.netD_att))}
        fill((1, 2))
        for _ in range(2):                                  # eager warm-up: builds every packed operand and plan
            ts.step(b, optimizer_step=True)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            ts.step(b, optimizer_step=True)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        calls[0] = 0
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            static = ts.step(b, optimizer_step=True)
        # back to the initial weights / statistics and fresh Adam moments, written IN PLACE (the graph holds the addresses)
        with torch.no_grad():
            for n, net in (("G", ts.netG), ("D_img", ts.netD_image), ("D_obj", ts.netD_object), ("D_att", ts.netD_att)):
                for k, v in net.state_dict().items():
                    v.copy_(sd0[n][k])
            for o in [ts.opt_G] + ts.opt_D:
                for st in o.state.values():
                    st["exp_avg"].zero_(); st["exp_avg_sq"].zero_()
                for t in o._steps.values():
                    t.zero_()
        ops.refresh_packs()                                 # weights were written behind the optimizer's back

        def replay(bb, seeds):
            fill(seeds)
            graph.replay()
            torch.cuda.synchronize()
            return static

        tol = dict(img_tol=1e-4, loss_tol=1e-4, cos_min=dict(G=0.97, D_img=0.999, D_obj=0.999, D_att=0.999)) \
            if precision == "fp32" else dict(img_tol=8e-2, loss_tol=5e-2, cos_min=dict(G=0.5, D_img=0.9, D_obj=0.9, D_att=0.9))
        run_synced_training(ts, batch, 64, states, 3, verbose=True, step_fn=replay, **tol)
    finally:
        ops.set_precision("fp32")


def test_ragged_batch_and_optimizer_steps():
    """3..9 objects per image incl. repeated steps with Adam: losses stay finite, weights move, step is deterministic"""
    batch = O.synth_batch(3, 64, None, 11)
    outs = []
    for _ in range(2):
        torch.manual_seed(0)
        ts = TrainStep(64, device="cuda")
        load_states(ts, O.make_states(64, 1))
        b = ts.to_device(batch)
        r = None
        for it in range(2):
            r = ts.step(b, optimizer_step=True, seeds=(5 + it, 50 + it))
        torch.cuda.synchronize()
        assert torch.isfinite(r["d_loss"]) and torch.isfinite(r["g_loss"])
        outs.append((float(r["d_loss"]), float(r["g_loss"]), r["out_g"][4].detach().clone()))
    assert outs[0][0] == outs[1][0] and outs[0][1] == outs[1][1] and torch.equal(outs[0][2], outs[1][2]), \
        "two identical runs must agree bit for bit (deterministic reductions)"


def test_full_size_properties_config2():
    """BASELINE config 2 shape (64x64, batch 32, 8 objects/image): size-independent checks — finite outputs, generated
    crops equal crops of generated images, per-image independence of the discriminator logits."""
    batch = O.synth_batch(32, 64, 8, 3)
    ts = TrainStep(64, device="cuda")
    b = ts.to_device(batch)
    with torch.no_grad():
        out = ts.generator(b, b["attribute"])["outputs"]
        from models.bilinear import crop_bbox_batch
        again = crop_bbox_batch(out[5], b["boxes"], b["obj_to_img"], 32)
        assert torch.equal(again, out[2])
        assert all(torch.isfinite(t).all() for t in out)
        ts.netD_image.eval()
        full = ts.netD_image(out[4])
        half = ts.netD_image(out[4][:16].contiguous())
        assert rel(half, full[:16]) < 1e-5


def test_graphed_train_step_layout_cache():
    """b200gan.graphed.GraphedTrainStep: an unseen layout runs one eager iteration and is captured; the same layout is then
    replayed from its graph with new batch VALUES (results equal the eager step on those values, bit for bit in fp32 — the
    kernels are deterministic); a different layout gets its own graph."""
    from b200gan.graphed import GraphedTrainStep, layout_key
    ops.set_precision("fp32")
    states = O.make_states(64, 0)

    def fresh():
        ts = TrainStep(64, device="cuda")
        load_states(ts, states)
        ts.netG.crop_encoder.eps_source = lambda o, z, d: torch.full((o, z), 0.25, device=d)    # no RNG in this comparison
        return ts

    a1 = O.synth_batch(2, 64, 4, 7)
    a2 = O.synth_batch(2, 64, 4, 8)            # same layout (2 images x 4 objects), other values
    c1 = O.synth_batch(3, 64, None, 9)         # another layout
    assert layout_key(a1) == layout_key(a2) != layout_key(c1)
    ts = fresh()
    gts = GraphedTrainStep(ts)
    seq = [a1, a2, c1, a1]
    got = []
    for hb in seq:
        r = gts.step(hb)
        torch.cuda.synchronize()
        got.append((float(r["d_loss"]), float(r["g_loss"]), r["out_g"][4].detach().clone()))
    assert gts.misses == 2 and gts.hits == 2 and all(it.graph is not None for it in gts.cache.values()), \
        [it.error for it in gts.cache.values()]
    ref_ts = fresh()
    for hb, g in zip(seq, got):
        r = ref_ts.step(ref_ts.to_device(hb), optimizer_step=True)
        torch.cuda.synchronize()
        assert abs(float(r["d_loss"]) - g[0]) <= 1e-6 * abs(g[0]) and abs(float(r["g_loss"]) - g[1]) <= 1e-6 * abs(g[1])
        assert rel(r["out_g"][4], g[2]) < 1e-6


def test_attribute_classifier_training_iterations():
    """evaluation/train_att_cls.py:196-258 (b200gan.att_cls.AttributeClassifierStep) against the oracle's attribute
    discriminator + torch's BCE + torch.optim.Adam: three iterations on 64x64 crops of 128x128 images, each compared from the
    same state (loss 1e-5, logits 1e-4, update direction cosine >= 0.999)."""
    from b200gan.att_cls import AttributeClassifierStep
    import torch.nn.functional as F
    ops.set_precision("fp32")
    states = O.make_states(64, 0)
    batch = O.synth_batch(3, 128, None, 21, sparse_attributes=True)
    st = AttributeClassifierStep(crop_size=64, device="cuda")
    st.net.load_state_dict(states["D_att"], strict=True)
    b = st.to_device(batch)
    pw = O.pos_weight_vector()
    for it in range(3):
        sd = {k: v.detach().cpu().clone().requires_grad_(O.is_parameter(k)) for k, v in st.net.state_dict().items()}
        before = {k: p.detach().cpu().clone() for k, p in st.net.named_parameters()}
        opt = torch.optim.Adam([v for v in sd.values() if v.requires_grad], lr=2e-4, betas=(0.5, 0.999))
        for k, p in st.net.named_parameters():
            so = st.opt.state.get(p)
            if so:
                opt.state[sd[k]] = dict(step=torch.tensor(float(so["step"])), exp_avg=so["exp_avg"].cpu().clone(),
                                        exp_avg_sq=so["exp_avg_sq"].cpu().clone())
        r = st.step(b)
        with torch.no_grad():
            crops = O.crop_bbox_batch(batch["imgs"], batch["boxes"], batch["obj_to_img"], 64)
        logits = O.attribute_discriminator(sd, crops)
        idx = batch["attribute"].sum(1).nonzero().view(-1)
        loss = F.binary_cross_entropy_with_logits(logits.index_select(0, idx), batch["attribute"].index_select(0, idx), pos_weight=pw)
        loss.backward()
        opt.step()
        assert abs(float(r["loss"]) - float(loss)) < 1e-5 * abs(float(loss)), (it, float(r["loss"]), float(loss))
        assert rel(r["logits"], logits) < 1e-4
        ua = torch.cat([(p.detach().cpu() - before[k]).reshape(-1) for k, p in st.net.named_parameters()]).double()
        ur = torch.cat([(sd[k].detach() - before[k]).reshape(-1) for k, _ in st.net.named_parameters()]).double()
        assert float(torch.nn.functional.cosine_similarity(ua, ur, dim=0)) > 0.999


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_forked_streams_bit_identical_to_single_stream(precision):
    """The fork / join concurrency of the iteration (three discriminators side by side, the G-step generator pass next to the
    D-step, deferred crop-encoder call, forked weight-gradient chains and data-gradient phases) only changes WHEN
    kernels run: two training iterations with every lever off and with every lever on must agree bit for bit — losses, all 11
    generator outputs, every parameter after the Adam updates, batch-norm running statistics and spectral-norm vectors.  Every
    kernel has fixed-order reductions, so any difference would be a missing dependency between streams."""
    import b200gan.step as step_mod
    batch = O.synth_batch(4, 64, None, 21)                      # ragged: 3..9 objects per image
    states = O.make_states(64, 3)
    flags = [(step_mod, "PARALLEL_D"), (step_mod, "OVERLAP_G2"), (step_mod, "DEFER_TAIL"), (ops, "SIDE_WGRAD"),
             (ops, "PARALLEL_PHASES"), (ops, "PER_STREAM_FORKS")]
    saved = [getattr(m, n) for m, n in flags]
    saved_mode = ops.FORKS
    runs = []
    ops.set_precision(precision)
    try:
        for on in (False, True, True):                          # the forked run twice: it must also agree with itself
            ops.FORKS = "always" if on else "off"               # (the default applies the levers only under graph capture)
            for m, n in flags:
                setattr(m, n, on)
            torch.manual_seed(0)
            ts = TrainStep(64, device="cuda")
            load_states(ts, states)
            b = ts.to_device(batch)
            res = None
            for it in range(2):
                res = ts.step(b, optimizer_step=True, seeds=(31 + it, 41 + it))
            torch.cuda.synchronize()
            snap = [res["d_loss"].clone(), res["g_loss"].clone()] + [t.clone() for t in res["out_g"]]
            for net in (ts.netG, ts.netD_image, ts.netD_object, ts.netD_att):
                snap += [v.detach().clone() for _, v in sorted(net.state_dict().items())]
            runs.append(snap)
    finally:
        ops.FORKS = saved_mode
        for (m, n), v in zip(flags, saved):
            setattr(m, n, v)
        ops.set_precision("fp32")
    for other in runs[1:]:
        assert len(other) == len(runs[0])
        for i, (a, r) in enumerate(zip(runs[0], other)):
            assert torch.equal(a, r), "tensor %d differs between the single-stream and the forked iteration" % i
