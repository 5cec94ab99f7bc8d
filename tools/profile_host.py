"""Host-side cost of one EAGER training iteration (no CUDA graph): cProfile of ts.step on the config-2 batch.
usage: python tools/profile_host.py [--ragged] > gpurun_out/host_profile.txt"""
import cProfile
import io
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "attribute-guided-image-generation-from-layout_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402
from b200gan import ops  # noqa: E402
from b200gan.step import TrainStep  # noqa: E402
from oracle import gan_oracle as O  # noqa: E402

ops.set_precision("bf16")
ts = TrainStep(64, device="cuda")
ts.netG.crop_encoder.eps_source = lambda o, z, d: torch.randn(o, z, device=d)
batches = [ts.to_device(O.synth_batch(32, 64, None if "--ragged" in sys.argv else 8, seed=s)) for s in range(4)]
for b in batches:
    ts.step(b)
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(8):
    ts.step(batches[i % 4])
torch.cuda.synchronize()
print("eager iteration: %.2f ms (host + device, %d launches per iteration)" % ((time.perf_counter() - t0) / 8 * 1e3,
                                                                                 0))
pr = cProfile.Profile()
pr.enable()
for i in range(4):
    ts.step(batches[i % 4])
torch.cuda.synchronize()
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(45)
print(s.getvalue())
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(60)
print(s.getvalue())
