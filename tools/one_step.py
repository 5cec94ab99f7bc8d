"""Minimal driver for ncu: two eager G+D steps at the bench workload (config 2) — first is warm-up."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "attribute-guided-image-generation-from-layout_b200"))
import torch
from b200gan import _lib, ops
from b200gan.step import TrainStep
from oracle import gan_oracle as O
size = int(os.environ.get("SIZE", "64")); batch = int(os.environ.get("BATCH", "32")); steps = int(os.environ.get("STEPS", "2"))
ops.set_precision(os.environ.get("PRECISION", "bf16"))
ts = TrainStep(size, device="cuda")
ts.netG.crop_encoder.eps_source = lambda o, z, d: torch.randn(o, z, device=d)
b = ts.to_device(O.synth_batch(batch, size, 8, 10))
for i in range(steps):
    torch.cuda.synchronize(); t = time.time()
    if i == steps - 1:
        torch.cuda.profiler.start()          # ncu --profile-from-start off: only the last step is captured
    n0 = _lib.K.launch_count()
    r = ts.step(b, optimizer_step=True)
    torch.cuda.synchronize()
    print("step %d: %.1f ms wall, %d library launches, d_loss %.4f g_loss %.4f" % (i, (time.time() - t) * 1e3, _lib.K.launch_count() - n0, float(r["d_loss"]), float(r["g_loss"])), flush=True)
torch.cuda.profiler.stop()
