#!/usr/bin/env python
"""bench.py — G+D train-step throughput (images/sec) of the B200-native path, the metric of BASELINE.json.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--size 64|128] [--batch B]
                  [--precision bf16|fp32] [--no-graph] [--no-cpu-baseline]

Workload at N=1: BASELINE.json configs[1] — 64x64 model (train64.py), batch 32, 8 objects/image, random-init weights,
synthetic VG-shaped layouts; one "step" = attribute estimation + D-step + G-step (forward, losses, backward) + the four
Adam updates.  For N>1 every rank runs the same per-GPU workload on its own shard (weak scaling) and gradients are
all-reduced over NCCL in buckets overlapped with backward (b200gan/ddp.py).

Prints ONE JSON line (see README / DESIGN.md for the keys).  `--impl reference` times the reference algorithm's CPU
path (the oracle port, oracle/gan_oracle.py — the reference itself is pure Python and does not exist on the GPU box)
on the host cores with a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "attribute-guided-image-generation-from-layout_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

OBJS_PER_IMAGE = 8

# The contract is ONE JSON line on stdout.  Libraries print there too (NCCL's version banner at communicator creation),
# so file descriptor 1 is pointed at stderr for the lifetime of the process and the result line goes to a private
# duplicate of the original stdout.
_RESULT_OUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)


def emit(line: dict):
    print(json.dumps(line), file=_RESULT_OUT, flush=True)


# ------------------------------------------------------------------------------------------------------------
# FLOP model (FlopCounterMode fit of the reference autograd): per G+D step
# ------------------------------------------------------------------------------------------------------------
def step_flops(size, n_images, n_objs, skip_dead=True):
    if size == 64:
        f = 73.3e9 * n_images + 65.7e9 * n_objs
        dead = 3.9e9 * n_images + 6.3e9 * n_objs      # D weight gradients of the G-step, skipped here
    else:
        f = 620.2e9 * n_images + 223.2e9 * n_objs
        dead = 3.9e9 * 4 * n_images + 6.3e9 * 4 * n_objs
    return f - (dead if skip_dead else 0.0)


# ------------------------------------------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md: the clocks line, sampled DURING the timed region)
# ------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx = float(parts[1])
            except ValueError:
                continue
            for n, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------
# reference arm: the oracle port on the host cores
# ------------------------------------------------------------------------------------------------------------
def cpu_reference_rate(size, n_images, steps, warmup, threads):
    from oracle import gan_oracle as O
    torch.set_num_threads(threads)
    states = O.make_states(size, 0)
    model = O.OracleModel(size, 0, states)
    batch = O.synth_batch(n_images, size, OBJS_PER_IMAGE, 3)
    opts = [torch.optim.Adam([v for v in st.values() if v.requires_grad], lr=2e-4, betas=(0.5, 0.999))
            for st in (model.G, model.D_img, model.D_obj, model.D_att)]

    def one():
        b = dict(batch)
        b["attribute_GT"] = b["attribute"].clone()
        nets = model.nets()
        with torch.no_grad():
            crops = O.crop_bbox_batch(b["imgs"], b["boxes"], b["obj_to_img"], model.obj_size)
        est = O.estimate_attributes(nets["att"](crops).detach(), b["attribute"])
        out = model.generator(b, est)
        d_loss, _ = O.d_step_loss(nets, b, out, model.pos_weight)
        model.zero_grad((model.D_img, model.D_obj, model.D_att))
        d_loss.backward()
        for o in opts[1:]:
            o.step()
        out = model.generator(b, est)
        g_loss, _ = O.g_step_loss(nets, b, out, model.pos_weight)
        model.zero_grad((model.G,))
        g_loss.backward()
        opts[0].step()
        return float(d_loss), float(g_loss)

    for _ in range(warmup):
        one()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    dt = time.perf_counter() - t0
    return n_images * steps / dt, dt / steps


def gpu_comparator(size, n_images, dev, steps=3):
    """The honest GPU comparator: the SAME algorithm as plain functional PyTorch — the oracle
    port, i.e. cuDNN convolutions / cuBLAS linears / ATen elementwise kernels, eager, fp32 storage with PyTorch's default cuDNN
    TF32 convolutions — on the same B200 and the same batch.  A reported baseline like cpu_baseline, never the product."""
    from oracle import gan_oracle as O
    states = O.make_states(size, 0)
    model = O.OracleModel(size, 0, states)
    for name in ("G", "D_img", "D_obj", "D_att"):
        st = getattr(model, name)
        setattr(model, name, {k: v.detach().to(dev).requires_grad_(v.requires_grad) for k, v in st.items()})
    model.pos_weight = model.pos_weight.to(dev)
    host = O.synth_batch(n_images, size, OBJS_PER_IMAGE, 3)
    batch = {k: (v if k == "obj_to_img" else v.to(dev)) for k, v in host.items()}
    opts = [torch.optim.Adam([v for v in st.values() if v.requires_grad], lr=2e-4, betas=(0.5, 0.999))
            for st in (model.G, model.D_img, model.D_obj, model.D_att)]

    def one():
        b = dict(batch)
        b["attribute_GT"] = b["attribute"].clone()
        nets = model.nets()
        with torch.no_grad():
            crops = O.crop_bbox_batch(b["imgs"], b["boxes"], b["obj_to_img"], model.obj_size)
        est = O.estimate_attributes(nets["att"](crops).detach(), b["attribute"])
        out = model.generator(b, est, eps=[torch.randn(b["objs"].shape[0], 64, device=dev) for _ in range(3)])
        d_loss, _ = O.d_step_loss(nets, b, out, model.pos_weight)
        model.zero_grad((model.D_img, model.D_obj, model.D_att))
        d_loss.backward()
        for o in opts[1:]:
            o.step()
        out = model.generator(b, est, eps=[torch.randn(b["objs"].shape[0], 64, device=dev) for _ in range(3)])
        g_loss, _ = O.g_step_loss(nets, b, out, model.pos_weight)
        model.zero_grad((model.G,))
        g_loss.backward()
        opts[0].step()
        return d_loss.detach(), g_loss.detach()

    one()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        dl, gl = one()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"what": "the same iteration as plain functional PyTorch (oracle port: cuDNN / cuBLAS / ATen kernels, eager, fp32 tensors, "
                    "PyTorch-default cuDNN TF32 convolutions) on the same B200, batch %d, %d objects/image" % (n_images, OBJS_PER_IMAGE),
            "ms_per_step": ms, "value": n_images / (ms / 1e3), "unit": "images/s", "steps": steps,
            "finite": bool(torch.isfinite(dl)) and bool(torch.isfinite(gl)), "kind": "port"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n = 2
    rate, sec = cpu_reference_rate(args.size, n, args.steps, args.warmup, threads)
    line = {
        "impl": "reference", "metric": "G+D train-step images/sec", "value": rate, "unit": "images/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, n_per_gpu=n, note="CPU sample"),
        "cpu_baseline": {"value": rate, "unit": "images/s", "cores": threads, "kind": "port",
                         "sample": "%d steps of the %dx%d G+D step (incl. Adam) at batch %d, %d objects/image, oracle port of the "
                                   "reference on %d host threads" % (args.steps, args.size, args.size, n, OBJS_PER_IMAGE, threads)},
        "e2e": {"value": rate, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def workload_config(args, n_per_gpu, note=""):
    return {"workload": "%dx%d model (train%d.py) G+D train step, batch %d per GPU, %d objects/image, random-init weights, "
                        "synthetic VG layouts%s" % (args.size, args.size, args.size, n_per_gpu, OBJS_PER_IMAGE,
                                                    (" [" + note + "]") if note else ""),
            "image_size": args.size, "batch_per_gpu": n_per_gpu, "objects_per_image": OBJS_PER_IMAGE,
            "parallelism": "dp%d" % args.gpus, "optimizer": "Adam (4 optimizers) inside the timed step",
            "l2": "per-step working set (activations + 61M-parameter weights/Adam state, > 1 GB) exceeds the 126 MB L2"}


# ------------------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------------------
def Iteration(ts, host, use_graph, world, rank, warm=2):
    """one training iteration on one batch layout, captured into a CUDA graph (b200gan/graphed.py — the product's own
    replay path): `upload()` copies the pinned host batch into the static device batch, `run()` replays"""
    from b200gan.graphed import CapturedIteration
    it = CapturedIteration(ts, host, use_graph, world > 1, warm=warm)
    if it.error and rank == 0:
        print("[bench] CUDA graph capture unavailable (%s); timing eagerly" % it.error, file=sys.stderr)
    return it


def timed(fn, steps, sync_all):
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = None
    for i in range(steps):
        out = fn(i)
    e1.record()
    sync_all()
    return e0.elapsed_time(e1) / steps, out


def max_over_ranks(vals, dev, world):
    import torch.distributed as dist
    t = torch.tensor(vals, device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t]


def make_step(size, precision, dev, world, graph):
    from b200gan import ops
    from b200gan.step import TrainStep
    ops.set_precision(precision)
    torch.manual_seed(1234)
    ts = TrainStep(size, device=dev, capturable=graph)
    if world > 1:
        ts.enable_data_parallel()  # the bucketed all-reduces issued from the autograd hooks are captured with the step
    # CropEncoder noise: drawn on the device (graph-safe Philox stream) instead of the reference's CPU RNG + H2D copy; the
    # distribution is the same, and the parity tests pin the CPU-RNG variant (tests/test_step_gpu.py)
    ts.netG.crop_encoder.eps_source = lambda o, z, d: torch.randn(o, z, device=d)
    return ts


def side_measurement(args, size, n_img, precision, steps, dev, world, rank, sync_all, note, label):
    """an extra workload measured in the same process (BASELINE configs other than the headline one): device-resident and
    end-to-end ms per iteration, max over ranks"""
    import gc
    from b200gan import _lib
    from oracle import gan_oracle as O
    try:
        ts = make_step(size, precision, dev, world, not args.no_graph)
        host = O.synth_batch(n_img, size, OBJS_PER_IMAGE, seed=40 + rank)
        it = Iteration(ts, host, not args.no_graph, world, rank, warm=2)
        for _ in range(2):
            it.run()
        ms, losses = timed(lambda i: it.run(), steps, sync_all)

        def e2e(i):
            it.upload()
            l = it.run()
            return torch.stack([l["d_loss"].reshape(()), l["g_loss"].reshape(())]).cpu()
        ms_e2e, hl = timed(e2e, steps, sync_all)
        ok = bool(torch.isfinite(hl).all())
        ms, ms_e2e = max_over_ranks([ms, ms_e2e], dev, world)
        mem = torch.cuda.max_memory_allocated(dev) / 2 ** 30
        res = {"workload": "%dx%d model, batch %d per GPU (global %d), %d objects/image, %s" %
                           (size, size, n_img, n_img * world, OBJS_PER_IMAGE, precision),
               "dtype": {"bf16": "bf16", "tf32": "tf32 (fp32 tensors, tcgen05 kind::tf32)", "fp32": "f32"}[precision],
               "ms_per_step": ms, "value": n_img * world / (ms / 1e3), "unit": "images/s", "steps": steps,
               "e2e": {"value": n_img * world / (ms_e2e / 1e3), "ms_per_step": ms_e2e,
                       "h2d_bytes_per_step": it.h2d_bytes * world, "d2h_bytes_per_step": 8 * world},
               "cuda_graph": it.graph is not None, "finite": ok, "peak_mem_gib": round(mem, 1),
               "step_tflops": step_flops(size, n_img, n_img * OBJS_PER_IMAGE) * world / (ms / 1e3) / 1e12}
        note("%s: %.2f ms/step" % (label, ms))
        del it, ts
    except Exception as e:          # an extra must never take the headline line down
        res = {"workload": label, "error": "%s: %s" % (type(e).__name__, str(e)[:300])}
        note("%s FAILED: %s" % (label, res["error"]))
    gc.collect()
    torch.cuda.empty_cache()
    return res


# ------------------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch.distributed as dist
    from b200gan import _lib, ops
    from oracle import gan_oracle as O       # synthetic batch generator + cpu_baseline leg only

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU for --impl b200 (there is no CPU fallback)"
    import __graft_entry__ as ge
    if not os.path.exists(ge.LIB):          # the in-tree library did not travel: local rank 0 builds it, the others wait
        if local == 0:
            ge.build()
        else:
            t_wait = time.time()
            while not os.path.exists(ge.LIB) and time.time() - t_wait < 600:
                time.sleep(2.0)
            time.sleep(2.0)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n_img = args.batch
    n_obj = n_img * OBJS_PER_IMAGE
    t_start = time.time()

    def note(msg):
        if os.environ.get("B200_BENCH_VERBOSE", "1") != "0":
            print("[bench r%d +%.1fs] %s" % (rank, time.time() - t_start, msg), file=sys.stderr, flush=True)

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ts = make_step(args.size, args.precision, dev, world, not args.no_graph)
    host = O.synth_batch(n_img, args.size, OBJS_PER_IMAGE, seed=10 + rank)
    # ---- eager warm-up + launch count -------------------------------------------------------------------------
    b0 = ts.to_device(host)
    for _ in range(max(2, args.warmup if args.no_graph else 2)):
        ts.step(b0, optimizer_step=True)
    torch.cuda.synchronize()
    launches_before = _lib.K.launch_count()
    t_e = time.perf_counter()
    ts.step(b0, optimizer_step=True)
    torch.cuda.synchronize()
    eager_ms = (time.perf_counter() - t_e) * 1e3
    launches_per_step = _lib.K.launch_count() - launches_before
    note("eager warm-up done (%d launches/step, %.1f ms eager)" % (launches_per_step, eager_ms))
    del b0
    main = Iteration(ts, host, not args.no_graph, world, rank, warm=1)
    for _ in range(args.warmup):
        main.run()
    torch.cuda.synchronize()
    note("graph captured" if main.graph is not None else "no graph: eager steps")

    # ---- device-resident timing (inputs already in HBM) ----------------------------------------------------------
    sampler = ClockSampler(local)
    traj = torch.zeros((args.steps, 2), device=dev)           # loss trajectory of the timed iterations (device-side copies)

    def dev_step(i):
        l = main.run()
        traj[i, 0].copy_(l["d_loss"]); traj[i, 1].copy_(l["g_loss"])
        return l["d_loss"], l["g_loss"]
    sync_all()
    if rank == 0:
        sampler.start()
    ms, losses = timed(dev_step, args.steps, sync_all)
    clocks = sampler.stop() if rank == 0 else None
    note("device-resident timing done: %.2f ms/step" % ms)
    assert all(torch.isfinite(l).all() for l in losses), "non-finite loss in the timed region"
    traj = traj.cpu()
    # the timed iterations TRAIN: the same batch is replayed, so the losses must move from iteration to iteration
    assert args.steps < 2 or (traj[1:] != traj[:-1]).any(dim=1).all(), \
        "losses repeat across timed iterations: the optimizer update is not reaching the next iteration"

    # ---- end-to-end: pinned host batch -> device every step, both losses read back every step ------------------------
    def e2e_step(i):
        main.upload()
        l = main.run()
        return torch.stack([l["d_loss"].reshape(()), l["g_loss"].reshape(())]).cpu()
    ms_e2e, host_losses = timed(e2e_step, args.steps, sync_all)
    note("end-to-end timing done: %.2f ms/step" % ms_e2e)
    assert torch.isfinite(host_losses).all()
    ms, ms_e2e = max_over_ranks([ms, ms_e2e], dev, world)

    # ---- end-to-end over ROTATING RAGGED layouts (3..9 objects per image, the loader's range vg_custom_mask.py:45):
    # every distinct layout is a different set of index plans and tensor shapes, i.e. its own captured graph ----------------
    ragged = None
    if not args.no_extras:
        try:
            n_lay = 4
            its = [Iteration(ts, O.synth_batch(n_img, args.size, None, seed=200 + 10 * rank + j), not args.no_graph, world,
                             rank, warm=1) for j in range(n_lay)]
            for it in its:
                it.run()

            def rag_step(i):
                it = its[i % n_lay]
                it.upload()
                l = it.run()
                return torch.stack([l["d_loss"].reshape(()), l["g_loss"].reshape(())]).cpu()
            k = max(args.steps, n_lay)
            ms_rag, hl = timed(rag_step, k, sync_all)
            (ms_rag,) = max_over_ranks([ms_rag], dev, world)
            objs = sum(it.n_objs for it in its) / n_lay
            # an UNSEEN layout runs eagerly once (plans built on the host, ~1700 launches issued from Python)
            unseen = []
            for j in range(3):      # three different unseen layouts: the first also pays the caching allocator's new blocks
                fresh = ts.to_device(O.synth_batch(n_img, args.size, None, seed=900 + 7 * j + rank))
                torch.cuda.synchronize()
                t_e = time.perf_counter()
                ts.step(fresh, optimizer_step=True)
                torch.cuda.synchronize()
                unseen.append((time.perf_counter() - t_e) * 1e3)
            first_ms = min(unseen)
            ragged = {"layouts": n_lay, "objects_per_image": "3..9 (mean %.2f)" % (objs / n_img), "steps": k,
                      "ms_per_step": ms_rag, "value": n_img * world / (ms_rag / 1e3), "unit": "images/s",
                      "objects_per_s": objs * world / (ms_rag / 1e3),
                      "fixed_layout_objects_per_s": n_obj * world / (ms_e2e / 1e3),
                      "first_seen_layout_eager_ms": first_ms, "first_seen_layout_eager_ms_all": unseen,
                      "note": "one captured graph per distinct layout (cache keyed by the per-image object counts); an unseen "
                              "layout pays one eager iteration"}
            note("ragged e2e: %.2f ms/step over %d layouts; unseen layout eager %.1f ms" % (ms_rag, n_lay, first_ms))
            del its, fresh
        except Exception as e:
            ragged = {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}

    # ---- data parallel: what the gradient exchange costs (the same captured iteration without the all-reduces) --------------
    dp = None
    if world > 1 and not args.no_extras:
        try:
            saved = (ts.ddp_d, ts.ddp_g)
            ts.ddp_d = ts.ddp_g = None
            solo = Iteration(ts, host, not args.no_graph, 1, rank, warm=1)
            solo.run()
            ms_solo, _ = timed(lambda i: solo.run(), args.steps, sync_all)
            ts.ddp_d, ts.ddp_g = saved
            (ms_solo,) = max_over_ranks([ms_solo], dev, world)
            dp = {"ms_per_step_without_allreduce": ms_solo, "exposed_allreduce_ms": ms - ms_solo,
                  "bucket_bytes_D": saved[0].bucket_bytes(), "bucket_bytes_G": saved[1].bucket_bytes(),
                  "note": "buckets in launch order; the LAST bucket of each backward (the networks' first layers, kept small) is "
                          "the one whose all-reduce cannot overlap backward compute"}
            note("without all-reduce: %.2f ms/step" % ms_solo)
            del solo
        except Exception as e:
            dp = {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}

    # ---- roofline of the dominant kernel family (tcgen05 gather-GEMMs), timed with CUDA events per launch ------
    roof = None
    if rank == 0:
        roof = kernel_roofline(ts, main.b, args)
        note("kernel roofline done")
    main_h2d, graph_used = main.h2d_bytes, main.graph is not None
    del main, ts
    import gc
    gc.collect()
    torch.cuda.empty_cache()

    # ---- the other BASELINE configurations, as extra keys of the same line ------------------------------------------------
    extras = {}
    if not args.no_extras and args.size == 64 and args.precision == "bf16":
        k = max(3, min(args.steps, 5))
        extras["config2_fp32_tensor_core"] = side_measurement(args, 64, n_img, "tf32", k, dev, world, rank, sync_all, note,
                                                              "config 2 fp32 half (tf32 tensor cores)")
        per_gpu = max(1, 128 // world)
        extras["config3_128x128_global_batch_128"] = side_measurement(args, 128, per_gpu, "bf16", 3, dev, world, rank, sync_all,
                                                                      note, "config 3 (128x128, global batch 128, strong scaling)")
        if rank == 0 and world == 1:
            try:
                extras["comparator_torch_cuda_eager"] = gpu_comparator(args.size, n_img, dev)
                note("torch eager comparator: %.1f ms/step" % extras["comparator_torch_cuda_eager"]["ms_per_step"])
            except Exception as e:
                extras["comparator_torch_cuda_eager"] = {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}
            import gc
            gc.collect()
            torch.cuda.empty_cache()
    ops.set_precision(args.precision)

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        rate, sec = cpu_reference_rate(args.size, 2, 2, 1, threads)
        cpu_base = {"value": rate, "unit": "images/s", "cores": threads, "kind": "port",
                    "sample": "1 warm-up + 2 timed %dx%d G+D steps (incl. Adam) at batch 2, %d objects/image, oracle port of the "
                              "reference on %d host threads" % (args.size, args.size, OBJS_PER_IMAGE, threads)}
    if rank == 0:
        total_imgs = n_img * world
        line = {
            "metric": "G+D train-step images/sec", "value": total_imgs / (ms / 1e3), "unit": "images/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": {"bf16": "bf16", "tf32": "tf32", "fp32": "f32"}[args.precision],
            "data": "synthetic", "config": workload_config(args, n_img),
            "e2e": {"value": total_imgs / (ms_e2e / 1e3), "unit": "images/s", "h2d_bytes_per_step": main_h2d * world,
                    "d2h_bytes_per_step": 8 * world, "ms_per_step": ms_e2e},
            "gpu_launches": launches_per_step * args.steps,
            "gpu_launches_per_step": launches_per_step,
            "cuda_graph": graph_used,
            "eager_ms_per_step": eager_ms,
            "clocks": clocks,
            "roofline": roof,
            "cpu_baseline": cpu_base,
            "step_tflops": step_flops(args.size, n_img, n_obj) * world / (ms / 1e3) / 1e12,
            "loss_trajectory": {"d_loss": [round(float(v), 5) for v in traj[:, 0][:8]],
                                "g_loss": [round(float(v), 5) for v in traj[:, 1][:8]],
                                "note": "first timed iterations on one repeated batch; they move because every iteration "
                                        "applies the four Adam updates and re-packs the GEMM operands"},
            "e2e_ragged_layouts": ragged,
            "data_parallel": dp,
            "extras": extras,
        }
        emit(line)
    if world > 1:
        # every rank leaves together and hard-exits: tearing down NCCL communicators that captured CUDA graphs still
        # reference can block at interpreter shutdown, and the result line is already out
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def kernel_roofline(ts, b, args):
    """Device time of every launch of the dominant kernel family (the tcgen05 gather-GEMMs) in one step, measured live
    with CUDA events on the launching stream: each distinct launch (descriptor + kernel path) met during one eager step
    is captured into a CUDA graph of R back-to-back launches on its real operands and replayed between two events, so
    the figure is kernel time without host launch gaps (events around single eager launches measure the host).  The
    algorithmic FLOPs of those launches (2*M*N*K of the convolution each computes) divided by the summed durations is
    `achieved`.  Peak = MEASURED_PEAKS.json bf16 sustained figure (kernels timed inside a long step), else the
    B200_PROFILING.md fallback."""
    from b200gan import _lib, ops
    real = _lib._K if _lib._K is not None else None
    if real is None:
        return None
    R = 4
    records = []          # (tc, flops, ms per launch)
    per_key = {}
    by_shape = {}
    orig_conv, orig_wgrad = real.conv_gemm, real.wgrad_gemm

    def flops_of(d):
        return 2.0 * d.B * d.Qh * d.Qw * d.Cout * d.Th * d.Tw * d.Cin

    def timed(fn, wgrad):
        tc_pos = 4 if wgrad else 5          # wgrad_gemm(desc, P, G, ws, splits, tc) / conv_gemm(desc, inp, wmat, bias, scale, out, tc, mask)

        def wrapper(desc, *a, **kw):
            tc = bool(kw["tc"] if "tc" in kw else a[tc_pos])
            fn(desc, *a, **kw)                                   # the step's own launch
            key = (wgrad, tc, bytes(desc)) + tuple(int(x) for x in a if isinstance(x, int))
            if key not in per_key:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    for _ in range(R):
                        fn(desc, *a, **kw)
                g.replay()
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                g.replay()
                e.record()
                e.synchronize()
                per_key[key] = s.elapsed_time(e) / R
                del g
            records.append((tc, flops_of(desc), per_key[key]))
            shape = (("wgrad" if wgrad else "conv"), tc, desc.B * desc.Qh * desc.Qw, desc.Cin, desc.Cout, desc.Th * desc.Tw,
                     desc.in_sy, desc.out_sy)
            a = by_shape.setdefault(shape, [0, 0.0, 0.0])
            a[0] += 1; a[1] += per_key[key]; a[2] += flops_of(desc)
        return wrapper

    real.conv_gemm = timed(orig_conv, False)
    real.wgrad_gemm = timed(orig_wgrad, True)
    ddp = (ts.ddp_d, ts.ddp_g)
    ts.ddp_d = ts.ddp_g = None        # rank 0 alone runs this step: no gradient exchange
    literal = [0.0, 0.0]              # (tcgen05, other) FLOPs of the same step in the reference's literal operation order
    try:
        ts.step(b, optimizer_step=False)
        torch.cuda.synchronize()
        if ops.POOLED_CONV:
            # the discriminator blocks run avg_pool2(conv3x3(h)) as one 4x4 stride-2 convolution (16/36 of the multiply-adds)
            # and the 1x1 shortcut after the pooling (1/4): the ALGORITHMIC work of the family is what the reference's
            # operation order costs — counted here by running the step once in that order (no timing)
            def counting(fn, wgrad):
                tc_pos = 4 if wgrad else 5

                def wrapper(desc, *a, **kw):
                    tc = bool(kw["tc"] if "tc" in kw else a[tc_pos])
                    literal[0 if tc else 1] += flops_of(desc)
                    fn(desc, *a, **kw)
                return wrapper
            real.conv_gemm, real.wgrad_gemm = counting(orig_conv, False), counting(orig_wgrad, True)
            ops.POOLED_CONV = False
            try:
                ts.step(b, optimizer_step=False)
                torch.cuda.synchronize()
            finally:
                ops.POOLED_CONV = True
    finally:
        real.conv_gemm, real.wgrad_gemm = orig_conv, orig_wgrad
        ts.ddp_d, ts.ddp_g = ddp
    if os.environ.get("B200_BENCH_SHAPES"):
        for shape, (n, ms_, fl) in sorted(by_shape.items(), key=lambda kv: -kv[1][1])[:60]:
            print("[shape] %-5s tc=%d M=%8d Cin=%4d Cout=%4d T=%2d s_in=%d s_out=%d | n=%3d %8.3f ms %7.1f TFLOP/s" %
                  (shape[0], shape[1], shape[2], shape[3], shape[4], shape[5], shape[6], shape[7], n, ms_, fl / ms_ / 1e9),
                  file=sys.stderr)
    tc_t = sum(t for tc, _, t in records if tc) / 1e3
    tc_f = sum(f for tc, f, _ in records if tc)
    simt_t = sum(t for tc, _, t in records if not tc) / 1e3
    simt_f = sum(f for tc, f, _ in records if not tc)
    try:
        mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        peak, src = float(mp["bf16_tflops_sustained"]), "measured (MEASURED_PEAKS.json bf16_tflops_sustained)"
    except Exception:
        peak, src = 1400.0, "fallback (B200_PROFILING.md sustained ~1.4 PFLOP/s)"
    if tc_t > 0:
        exe_f, alg_f, fam_t = tc_f, (literal[0] if literal[0] > 0 else tc_f), tc_t
        name = "conv_gemm_tc_kernel + wgrad_gemm_tc_kernel (tcgen05 gather-GEMMs)"
    else:
        exe_f, alg_f, fam_t = simt_f, (literal[1] if literal[1] > 0 else simt_f), max(simt_t, 1e-9)
        name = "conv_gemm_f32_kernel + wgrad_gemm_f32_kernel (fp32 CUDA-core gather-GEMMs)"
    ach = alg_f / fam_t / 1e12
    traffic, traffic_src = None, None
    try:        # measured DRAM bytes of the family's launches in one iteration: the committed ncu metrics pass of this build
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic_r02.json")))
        if args.size == 64 and args.batch == 32 and args.precision == "bf16":
            traffic, traffic_src = float(tr["dram_bytes"]), tr["source"]
    except Exception:
        pass
    return {"bound": "tensor", "kernel": name, "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
            "peak_source": src, "traffic": traffic, "traffic_unit": "bytes of DRAM traffic per iteration, summed over the family's launches (ncu dram__bytes_read.sum + dram__bytes_write.sum)",
            "traffic_source": traffic_src, "launches": len([1 for r in records if r[0]]) if tc_t > 0 else len(records),
            "distinct_launch_shapes": len(per_key),
            "timing": "CUDA events around a graph of %d back-to-back launches per distinct launch shape" % R,
            "kernel_time_ms_per_step": (tc_t if tc_t > 0 else simt_t) * 1e3,
            "other_gemm_ms_per_step": (simt_t if tc_t > 0 else 0.0) * 1e3,
            "algorithmic_tflop_per_step": alg_f / 1e12,
            "executed_tflop_per_step": exe_f / 1e12, "executed_tflops": exe_f / fam_t / 1e12,
            "executed_frac": exe_f / fam_t / 1e12 / peak,
            "flops_counted": "achieved = ALGORITHMIC FLOPs (2*M*N*K of the family's launches in the reference's literal operation "
                             "order, counted by running the step once in that order) / measured time of the launches as run; "
                             "executed_* = 2*M*N*K of the launches as run" + (
                "; they differ because discriminator blocks run avg_pool2(conv3x3) as one 4x4 stride-2 convolution (16/36 of the "
                "multiply-adds) and the 1x1 shortcut after the pooling" if ops.POOLED_CONV else " (identical here)")}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--size", type=int, default=64, choices=[64, 128])
    ap.add_argument("--batch", type=int, default=32, help="images per GPU")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "tf32", "fp32"])
    ap.add_argument("--no-extras", action="store_true", help="headline workload only (no ragged / tf32 / 128x128 extras)")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
