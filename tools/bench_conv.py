"""GPU microbenchmark of the tcgen05 gather-GEMMs on the shapes of the G+D step (config 2), per kernel path.

Every (shape, op, mode) is captured into a CUDA graph of R back-to-back launches and replayed, so the figure is device
time per launch without host gaps.  Modes: cpasync / im2col (one tile per CTA) / persist / halo (shifted window).
  python tools/bench_conv.py [fwd|dgrad|wgrad|all] [shape-filter]
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "attribute-guided-image-generation-from-layout_b200"))
import torch
from b200gan import _lib, ops

# (Cx, Cy, k, s, p, H, N)  — dominant shapes of the 64x64 step at batch 32 / 256 objects
SHAPES = [
    (64, 64, 1, 1, 0, 32, 768),      # 1x1 shortcuts / im2col-packed 3-channel convolutions
    (64, 128, 1, 1, 0, 32, 768),
    (128, 256, 1, 1, 0, 16, 768),
    (64, 64, 3, 1, 1, 32, 768),      # D_obj / D_att blocks at 32x32 (3 batched calls)
    (64, 64, 3, 1, 1, 64, 256),      # D_img block at 64x64, SPADE_3 shared conv
    (64, 128, 3, 1, 1, 32, 768),
    (128, 128, 3, 1, 1, 16, 768),
    (128, 128, 3, 1, 1, 32, 96),
    (128, 256, 3, 1, 1, 16, 384),
    (256, 256, 3, 1, 1, 8, 768),
    (256, 512, 3, 1, 1, 8, 768),
    (512, 512, 3, 1, 1, 4, 768),
    (64, 128, 4, 2, 1, 66, 768),     # LayoutEncoder c2
    (128, 256, 4, 2, 1, 33, 768),
    (256, 512, 4, 2, 1, 16, 768),
    (64, 128, 4, 2, 1, 32, 768),     # CropEncoder
    (512, 512, 5, 1, 2, 8, 768),     # ConvLSTM layer 0 input-to-gate conv over all objects
    (128, 512, 5, 1, 2, 8, 96),      # ConvLSTM hidden-to-gate, one time step
    (64, 256, 5, 1, 2, 8, 96),
    (64, 3, 7, 1, 3, 64, 96),        # decoder c4
    (192, 256, 3, 1, 1, 8, 96),
    (64, 64, 3, 1, 1, 8, 96),        # G residual blocks
    (256, 64, 5, 1, 2, 8, 96),
    (512, 128, 5, 1, 2, 8, 96),
    (256, 256, 4, 2, 1, 16, 96),
]
MODES = {"cpasync": (False, 0, False), "im2col": (True, 0, False), "persist": (True, 1, False),
         "persist2": (True, 2, False), "halo": (True, 1, True)}


def timeit(fn, R=10, reps=5):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn()
        with torch.cuda.graph(g, stream=s):
            for _ in range(R):
                fn()
    torch.cuda.synchronize()
    g.replay(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / R)
    return best * 1e3      # us


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    filt = sys.argv[2] if len(sys.argv) > 2 else ""
    modes = os.environ.get("MODES", "im2col,persist,persist2").split(",")
    ops.set_precision("bf16")
    K = _lib.K
    print("%-34s %-6s " % ("shape (Cx,Cy,k,s,p,H,N)", "op") + " ".join("%16s" % m for m in modes))
    for shp in SHAPES:
        if filt and filt not in str(shp).replace(" ", ""):
            continue
        Cx, Cy, k, s, p, H, N = shp
        geom = ops.ConvGeom(Cx, Cy, k, k, s, p)
        Hy = geom.out_hw(H, H)[0]
        g = torch.Generator().manual_seed(1)
        x = torch.randn(N, H, H, Cx, generator=g).cuda().bfloat16()
        dy = torch.randn(N, Hy, Hy, Cy, generator=g).cuda().bfloat16()
        w = (torch.randn(Cy, Cx, k, k, generator=g) / (Cx * k * k) ** 0.5).cuda()
        dw = torch.empty_like(w)
        flop = 2.0 * N * Hy * Hy * Cy * Cx * k * k
        for op in ("fwd", "dgrad", "wgrad"):
            if which not in ("all", op):
                continue
            if op == "fwd" and Cx % 64: continue
            if op == "dgrad" and Cy % 64: continue
            if op == "wgrad" and (Cx % 64 or Cy % 64): continue
            cells = []
            for m in modes:
                a, b, c = MODES[m]
                pa, pb, pc = K.conv_tc_set_im2col(a), K.conv_tc_set_persistent(b), K.conv_tc_set_halo(c)
                try:
                    packs = ops.WeightPacks()
                    if op == "fwd":
                        fn = lambda: ops.conv_forward(geom, packs, w, x, "cl", "cl")
                    elif op == "dgrad":
                        fn = lambda: ops.conv_dgrad(geom, packs, w, dy, "cl", (H, H), "cl")
                    else:
                        if m in ("persist", "persist2", "halo"):
                            cells.append("%16s" % "-"); continue
                        fn = lambda: ops.conv_wgrad(geom, x, "cl", dy, "cl", dw)
                    us = timeit(fn)
                    cells.append("%7.1fus %5.0fTF" % (us, flop / us * 1e-6))
                finally:
                    K.conv_tc_set_im2col(pa); K.conv_tc_set_persistent(pb); K.conv_tc_set_halo(pc)
            print("%-34s %-6s " % (str(shp).replace(" ", ""), op) + " ".join(cells), flush=True)


if __name__ == "__main__":
    main()
