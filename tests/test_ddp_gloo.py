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
    bk = GradBucketer(params, bucket_bytes=256)
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
