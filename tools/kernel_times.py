"""Per-kernel device time of one eager G+D step (config 2) from the CUPTI activity trace (torch.profiler): kernel durations
as they run inside the step (warm caches, no serialisation), summed per kernel name.  Cheaper than an ncu launch list;
used to check single-kernel changes.   python tools/kernel_times.py [top_n]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "attribute-guided-image-generation-from-layout_b200"))
import torch
from torch.profiler import profile, ProfilerActivity
from b200gan import ops
from b200gan.step import TrainStep
from oracle import gan_oracle as O

size = int(os.environ.get("SIZE", "64")); batch = int(os.environ.get("BATCH", "32"))
ops.set_precision(os.environ.get("PRECISION", "bf16"))
ts = TrainStep(size, device="cuda")
ts.netG.crop_encoder.eps_source = lambda o, z, d: torch.randn(o, z, device=d)
b = ts.to_device(O.synth_batch(batch, size, 8, 10))
for _ in range(2):
    ts.step(b, optimizer_step=True)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    ts.step(b, optimizer_step=True)
    torch.cuda.synchronize()
rows = []
for e in prof.key_averages():
    t = getattr(e, "device_time_total", None)
    if t is None:
        t = getattr(e, "cuda_time_total", 0.0)
    if t > 0:
        rows.append((t, e.count, e.key))
rows.sort(reverse=True)
tot = sum(r[0] for r in rows)
print("total kernel time %.2f ms over %d launches" % (tot / 1e3, sum(r[1] for r in rows)))
for t, n, k in rows[: int(sys.argv[1]) if len(sys.argv) > 1 else 60]:
    print("%8.3f ms %5d  %5.1f%%  %s" % (t / 1e3, n, 100 * t / tot, k[:110].replace("b200::", "")))
