"""GPU diagnostic: per-entry-point and per-conv-shape CUDA-event timing of one eager G+D step (not a bench number)."""
import argparse, json, os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "attribute-guided-image-generation-from-layout_b200"))
import torch
from b200gan import _lib, ops
from b200gan.step import TrainStep
from oracle import gan_oracle as O

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=64); ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--precision", default="bf16"); ap.add_argument("--out", default="gpurun_out/profile_step.json")
a = ap.parse_args()
ops.set_precision(a.precision)
ts = TrainStep(a.size, device="cuda")
ts.netG.crop_encoder.eps_source = lambda o, z, d: torch.randn(o, z, device=d)
b = ts.to_device(O.synth_batch(a.batch, a.size, 8, 10))
for _ in range(2):
    ts.step(b, optimizer_step=True)
torch.cuda.synchronize()
K = _lib._K
if os.environ.get("B200_NO_PERSIST"):
    K.conv_tc_set_persistent(False)
recs = []
names = [n for n in dir(K) if not n.startswith("_") and callable(getattr(K, n)) and n not in ("launch_count", "version", "bn_chunks", "conv_tc_ntile")]
orig = {n: getattr(K, n) for n in names}
def wrap(n, fn):
    def w(*args, **kw):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); r = fn(*args, **kw); e.record()
        info = None
        if n in ("conv_gemm", "wgrad_gemm"):
            d = args[0]
            info = dict(tc=bool(args[-1]), M=d.B * d.Qh * d.Qw, Cin=d.Cin, Cout=d.Cout, T=d.Th * d.Tw, B=d.B, Qh=d.Qh,
                        flops=2.0 * d.B * d.Qh * d.Qw * d.Cout * d.Th * d.Tw * d.Cin, splits=(args[4] if n == "wgrad_gemm" else 0))
        recs.append((n, s, e, info))
        return r
    return w
for n in names:
    setattr(K, n, wrap(n, orig[n]))
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record(); ts.step(b, optimizer_step=True); t1.record()
torch.cuda.synchronize()
for n in names:
    setattr(K, n, orig[n])
tot = collections.defaultdict(lambda: [0, 0.0])
shapes = collections.defaultdict(lambda: [0, 0.0, 0.0])
for n, s, e, info in recs:
    ms = s.elapsed_time(e)
    key = n if info is None else "%s[%s]" % (n, "tc" if info["tc"] else "f32")
    tot[key][0] += 1; tot[key][1] += ms
    if info is not None:
        k2 = (n, info["tc"], info["M"], info["Cin"], info["Cout"], info["T"], info["splits"])
        shapes[k2][0] += 1; shapes[k2][1] += ms; shapes[k2][2] += info["flops"]
print("step (with event overhead): %.1f ms; sum of kernel-call spans: %.1f ms; calls %d" % (t0.elapsed_time(t1), sum(v[1] for v in tot.values()), len(recs)))
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print("%-28s calls %5d  %8.2f ms" % (k, v[0], v[1]))
print("--- conv shapes by time")
rows = sorted(shapes.items(), key=lambda kv: -kv[1][1])
for k, v in rows[:45]:
    print("%-10s tc=%d M=%8d Cin=%4d Cout=%4d T=%2d splits=%2d | n=%3d  %7.2f ms  %7.1f TFLOP/s" % (k[0], k[1], k[2], k[3], k[4], k[5], k[6], v[0], v[1], v[2] / v[1] / 1e9))
json.dump({"totals": {k: v for k, v in tot.items()}, "shapes": [[list(k), v] for k, v in rows]}, open(a.out, "w"))
