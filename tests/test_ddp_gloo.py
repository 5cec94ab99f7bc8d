"""not gpu: the data-parallel plumbing (bucketed, hook-driven gradient all-reduce) on world_size=2 gloo processes."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "attribute-guided-image-generation-from-layout_b200")


def _worker(rank, world, port, out):
    sys.path.insert(0, PKG)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from b200gan.ddp import GradBucketer, broadcast_module, shard_images
    torch.manual_seed(rank)             # different init per rank: broadcast must fix it
    net = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.ReLU(), torch.nn.Linear(32, 32), torch.nn.ReLU(),
                              torch.nn.Linear(32, 4), torch.nn.Linear(4, 4))
    broadcast_module(net)
    unused = torch.nn.Linear(3, 3)     # a parameter that never receives a gradient
    params = list(net.parameters()) + list(unused.parameters())
    bk = GradBucketer(params, bucket_bytes=256, tail_bytes=200)
    assert len(bk.buckets) >= 3
    g = torch.Generator().manual_seed(42)
    x = torch.randn(8, 16, generator=g)
    lo, hi = shard_images(8, rank, world)
    for it in range(2):                 # twice: re-arming works
        net.zero_grad(set_to_none=True)
        bk.arm()
        net(x[lo:hi]).pow(2).mean().backward()
        bk.finish()
    grads = [p.grad.clone() for p in net.parameters()]
    if rank == 0:
        ref = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.ReLU(), torch.nn.Linear(32, 32), torch.nn.ReLU(),
                                  torch.nn.Linear(32, 4), torch.nn.Linear(4, 4))
        ref.load_state_dict(net.state_dict())
        total = [torch.zeros_like(p) for p in ref.parameters()]
        for r in range(world):
            a, b = shard_images(8, r, world)
            ref.zero_grad(set_to_none=True)
            ref(x[a:b]).pow(2).mean().backward()
            for t, p in zip(total, ref.parameters()):
                t += p.grad / world
        err = max(float((g_ - t).abs().max()) for g_, t in zip(grads, total))
        torch.save({"err": err, "unused_none": all(p.grad is None for p in unused.parameters())}, out)
    dist.barrier()
    dist.destroy_process_group()


def test_bucketed_allreduce_world2(tmp_path):
    out = str(tmp_path / "res.pt")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    res = torch.load(out)
    assert res["err"] < 1e-6, res
    assert res["unused_none"]


def test_shard_images_partition():
    sys.path.insert(0, PKG)
    from b200gan.ddp import shard_images
    for n, w in ((128, 8), (10, 4), (3, 8), (32, 1)):
        spans = [shard_images(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))


def _step_worker(rank, world, port, out_dir, backend="gloo"):
    """one rank of a data-parallel G+D step: its shard of the batch, bucketed all-reduce.  gloo: the CPU emulation of the
    kernels; nccl: the CUDA kernels on device `rank` (tests/test_ddp_gpu.py)"""
    for p in (ROOT, PKG, os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    from b200gan import _lib
    if backend == "gloo":
        dist.init_process_group("gloo", rank=rank, world_size=world)
        torch.set_num_threads(max(1, (os.cpu_count() or 2) // world))
        from abi_emul import EmulKernels
        _lib.K = EmulKernels()
        device = "cpu"
    else:
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
        device = "cuda:%d" % rank
    from b200gan.step import TrainStep
    from helpers import load_states
    from oracle import gan_oracle as O
    states = O.make_states(64, 0 if rank == 0 else 5)      # rank 1 starts from OTHER weights: the broadcast must fix that
    ts = TrainStep(64, device=device, optimizer="torch" if backend == "gloo" else "b200")
    load_states(ts, states)
    ts.enable_data_parallel(bucket_bytes=8 << 20)
    batch = O.synth_batch(1, 64, 4, 20 + rank)              # one image per rank (objects travel with their image)
    res = ts.step(ts.to_device(batch), optimizer_step=False, seeds=(123 + rank, 223 + rank))
    grads = {n: {k: p.grad.detach().cpu().clone() for k, p in net.named_parameters()} for n, net in
             (("G", ts.netG), ("D_img", ts.netD_image), ("D_obj", ts.netD_object), ("D_att", ts.netD_att))}
    torch.save(dict(grads=grads, d_loss=float(res["d_loss"]), g_loss=float(res["g_loss"])), os.path.join(out_dir, "rank%d.pt" % rank))
    dist.barrier()
    dist.destroy_process_group()


def check_against_shard_oracles(tmp_path):
    _check(tmp_path)


def test_data_parallel_step_world2_matches_mean_of_shard_oracles(tmp_path):
    """the data-parallel target is mean_r(grad_ref(shard_r)) — the CPU oracle run on each shard — for the D
    gradients after d_loss.backward() and the G gradients after g_loss.backward(); both ranks must hold the same reduced
    gradients.  (The G+D step itself, not a toy network: every parameter of the four networks goes through the buckets.)"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import oracle_step, rel
    from oracle import gan_oracle as O
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_step_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    _check(tmp_path)


def _check(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import oracle_step, rel
    from oracle import gan_oracle as O
    r0, r1 = torch.load(str(tmp_path / "rank0.pt")), torch.load(str(tmp_path / "rank1.pt"))
    states = O.make_states(64, 0)
    refs = [oracle_step(O.OracleModel(64, 0, states), O.synth_batch(1, 64, 4, 20 + r), seeds=(123 + r, 223 + r)) for r in range(2)]
    assert abs(r0["d_loss"] - float(refs[0]["d_loss"])) < 1e-4 * abs(float(refs[0]["d_loss"]))
    assert abs(r1["g_loss"] - float(refs[1]["g_loss"])) < 1e-4 * abs(float(refs[1]["g_loss"]))
    cosf = torch.nn.functional.cosine_similarity
    for net in ("G", "D_img", "D_obj", "D_att"):
        a, b, m = [], [], []
        for k, g0 in r0["grads"][net].items():
            assert torch.equal(g0, r1["grads"][net][k]), (net, k)            # identical on both ranks after the all-reduce
            want = 0.5 * ((refs[0]["g_grads"][k] + refs[1]["g_grads"][k]) if net == "G" else
                          (refs[0]["d_grads"][net][k] + refs[1]["d_grads"][net][k]))
            a.append(g0.reshape(-1).double())
            m.append(want.reshape(-1).double())
        c = float(cosf(torch.cat(a), torch.cat(m), dim=0))
        assert c > (0.9999 if net != "G" else 0.999), (net, c)
        assert rel(torch.cat(a), torch.cat(m)) < (1e-3 if net != "G" else 5e-2), net


def _shard_batch(batch, lo, hi):
    """images [lo, hi) of a batch and their objects (obj_to_img re-based), as a loader would hand them to one rank"""
    o2i = batch["obj_to_img"]
    rows = ((o2i >= lo) & (o2i < hi)).nonzero().view(-1)
    out = {}
    for k, v in batch.items():
        if k == "imgs":
            out[k] = v[lo:hi].clone()
        elif k == "obj_to_img":
            out[k] = (o2i[rows] - lo).clone()
        else:
            out[k] = v[rows].clone()
    return out, rows


def _syncbn_worker(rank, world, port, out_dir):
    for p in (ROOT, PKG, os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(max(1, (os.cpu_count() or 2) // world))
    from abi_emul import EmulKernels
    from b200gan import _lib
    from b200gan.ddp import shard_images
    _lib.K = EmulKernels()
    from b200gan.step import TrainStep
    from helpers import load_states
    from oracle import gan_oracle as O
    states = O.make_states(64, 0)
    ts = TrainStep(64, device="cpu", optimizer="torch")
    load_states(ts, states)
    ts.enable_data_parallel(bucket_bytes=8 << 20, sync_bn=True)
    full = O.synth_batch(3, 64, None, 31, sparse_attributes=True)          # ragged: 3..9 objects per image
    lo, hi = shard_images(3, rank, world)                                   # rank 0: images 0-1, rank 1: image 2
    batch, rows = _shard_batch(full, lo, hi)
    # the CropEncoder noise of the GLOBAL batch (one draw of (O, z) per generator forward feeds z_rec), this rank's rows
    eps = {}
    for seed in (123, 124):
        torch.manual_seed(seed)
        eps[seed] = torch.randn(full["objs"].shape[0], 64)[rows]
    state = {"seed": 123, "call": 0}

    def eps_source(o, z, dev):
        state["call"] += 1
        return eps[state["seed"]].clone() if o == rows.numel() else torch.zeros(o, z)   # later calls: only mu is used

    ts.netG.crop_encoder.eps_source = eps_source
    b = ts.to_device(batch)
    # D-step and G-step inside ts.step use the two generator forwards in order: switch the noise set between them
    orig_generator = ts.generator
    calls = {"n": 0}

    def generator(bb, est):
        state["seed"] = 123 if calls["n"] == 0 else 124
        calls["n"] += 1
        return orig_generator(bb, est)
    ts.generator = generator
    res = ts.step(b, optimizer_step=False)
    grads = {n: {k: p.grad.detach().clone() for k, p in net.named_parameters()} for n, net in
             (("G", ts.netG), ("D_img", ts.netD_image), ("D_obj", ts.netD_object), ("D_att", ts.netD_att))}
    stats = {k: v.clone() for k, v in ts.netG.state_dict().items() if "running_" in k}
    torch.save(dict(grads=grads, stats=stats, d_loss=float(res["d_loss"]), g_loss=float(res["g_loss"])),
               os.path.join(out_dir, "rank%d.pt" % rank))
    dist.barrier()
    dist.destroy_process_group()


def test_sync_bn_data_parallel_equals_the_global_batch_step(tmp_path):
    """with synchronised batch statistics and count-weighted losses the 2-rank step (2 + 1 images,
    ragged objects, sparse attributes) reproduces the SINGLE-PROCESS oracle step on the concatenated batch: gradients of all
    four networks, BN running statistics, and the (count-weighted, rank-averaged) losses."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import rel
    from oracle import gan_oracle as O
    port = 35500 + (os.getpid() % 2000)
    mp.spawn(_syncbn_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = torch.load(str(tmp_path / "rank0.pt")), torch.load(str(tmp_path / "rank1.pt"))
    states = O.make_states(64, 0)
    model = O.OracleModel(64, 0, states)
    full = O.synth_batch(3, 64, None, 31, sparse_attributes=True)
    b = dict(full)
    b["attribute_GT"] = b["attribute"].clone()
    nets = model.nets()
    with torch.no_grad():
        crops = O.crop_bbox_batch(b["imgs"], b["boxes"], b["obj_to_img"], 32)
    est = O.estimate_attributes(nets["att"](crops).detach(), b["attribute"])
    n_obj = b["objs"].shape[0]
    torch.manual_seed(123)
    e_d = [torch.randn(n_obj, 64), torch.zeros(n_obj, 64), torch.zeros(n_obj, 64)]
    out_d = model.generator(b, est, eps=e_d)
    d_loss, _ = O.d_step_loss(nets, b, out_d, model.pos_weight)
    model.zero_grad((model.D_img, model.D_obj, model.D_att))
    d_loss.backward()
    d_grads = {n: {k: v.grad.clone() for k, v in st.items() if v.requires_grad} for n, st in
               (("D_img", model.D_img), ("D_obj", model.D_obj), ("D_att", model.D_att))}
    torch.manual_seed(124)
    e_g = [torch.randn(n_obj, 64), torch.zeros(n_obj, 64), torch.zeros(n_obj, 64)]
    out_g = model.generator(b, est, eps=e_g)
    g_loss, _ = O.g_step_loss(nets, b, out_g, model.pos_weight)
    model.zero_grad((model.G,))
    g_loss.backward()
    # losses: each rank reports world * share-weighted terms; their mean over the ranks is the global loss
    assert abs(0.5 * (r0["d_loss"] + r1["d_loss"]) - float(d_loss)) < 1e-4 * abs(float(d_loss))
    assert abs(0.5 * (r0["g_loss"] + r1["g_loss"]) - float(g_loss)) < 1e-4 * abs(float(g_loss))
    cosf = torch.nn.functional.cosine_similarity
    for net in ("G", "D_img", "D_obj", "D_att"):
        a, m = [], []
        for k, g0 in r0["grads"][net].items():
            assert torch.equal(g0, r1["grads"][net][k]), (net, k)
            want = model.G[k].grad if net == "G" else d_grads[net][k]
            a.append(g0.reshape(-1).double())
            m.append(want.reshape(-1).double())
        c = float(cosf(torch.cat(a), torch.cat(m), dim=0))
        assert c > (0.9999 if net != "G" else 0.999), (net, c)
        assert rel(torch.cat(a), torch.cat(m)) < (1e-3 if net != "G" else 5e-2), net
    for k, v in r0["stats"].items():                       # running statistics follow the GLOBAL batch on every rank
        assert rel(v, model.G[k]) < 1e-4 or float((v - model.G[k]).abs().max()) < 1e-6, k
        assert torch.equal(v, r1["stats"][k]), k
