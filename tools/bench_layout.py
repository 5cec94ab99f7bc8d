"""BASELINE config 5 — layout microbenchmark at 128x128, N = 32 images, 8..30 boxes per image: device rasteriser, the
embedding (x) mask broadcast of LayoutEncoder c0 (rank-1 form), masks_to_layout (D = 128, 16x16 masks; forward and backward
w.r.t. embeddings + masks) and the 64x64 box crops forward + backward.  Device time per
launch (CUDA graph of R launches) and achieved GB/s on the ALGORITHMIC bytes of SURVEY.md §8(d) against the measured HBM
peak (MEASURED_PEAKS.json hbm_gbs).   python tools/bench_layout.py"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "attribute-guided-image-generation-from-layout_b200"))
import torch
from b200gan import _lib, layout, ops
from oracle import gan_oracle as O          # synthetic batch generator only

try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    PEAK = 6551.0


def timeit(fn, R=10, reps=5):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph(); s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn()
        with torch.cuda.graph(g, stream=s):
            for _ in range(R):
                fn()
    torch.cuda.synchronize(); g.replay(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / R)
    return best * 1e3


N, H, S, C = 32, 128, 64, 64
print("boxes/img  %-28s %-28s %-28s %-28s %-28s %-28s" % ("rasterise", "broadcast (bf16 out)", "crop fwd", "crop bwd",
                                                                  "masks_to_layout fwd", "masks_to_layout bwd"))
D, M = 128, 16
for per in (8, 12, 16, 20, 24, 30):
    b = O.synth_batch(N, H, per, seed=per)
    boxes, o2i = b["boxes"].cuda(), b["obj_to_img"]
    On = boxes.shape[0]
    masks = layout.rasterize_boxes(boxes, H, H)
    assert torch.equal(masks.cpu(), b["masks"])
    v = torch.randn(On, C, device="cuda")
    feats = torch.randn(N, 3, H, H, device="cuda")
    crops = ops.crop_bbox_batch(feats.detach(), boxes, o2i, S)
    gy = torch.randn_like(crops)
    plan = ops.get_plan(o2i, N, "cuda")
    wgt = ops.crop_weights(S, "cuda")
    vecs = torch.randn(On, D, device="cuda")
    m16 = torch.rand(On, M, M, device="cuda")
    lay = layout.masks_to_layout(vecs, boxes, m16, o2i, H, N=N)
    glay = torch.randn_like(lay)
    linx = layout._linspace01(H, "cuda")
    cells = []
    for fn, nbytes in ((lambda: layout.rasterize_boxes(boxes, H, H), 16 * On + 4 * On * H * H),
                       (lambda: _lib.K.mask_outer_fwd(v, masks, On, H, H, C, torch.bfloat16), 4 * On * H * H + 4 * On * C + 2 * On * C * (H + 2) ** 2),
                       (lambda: ops.crop_bbox_batch(feats.detach(), boxes, o2i, S), 4 * (On * 3 * S * S + N * 3 * H * H)),
                       (lambda: _lib.K.crop_bwd(gy, boxes, plan.img_box_start, plan.box_order, wgt, wgt, N, H, H), 4 * (On * 3 * S * S + N * 3 * H * H)),
                       (lambda: _lib.K.m2l_fwd(vecs, boxes, m16, plan.img_box_start, plan.box_order, linx, linx, N),
                        4 * (N * D * H * H + On * D + On * M * M)),
                       (lambda: _lib.K.m2l_bwd(glay, vecs, boxes, m16, plan.box_to_img, linx, linx), 4 * (N * D * H * H + 2 * On * D + 2 * On * M * M))):
        us = timeit(fn)
        gbs = nbytes / us * 1e-3
        cells.append("%7.1f us %6.0f GB/s %4.1f%%" % (us, gbs, 100 * gbs / PEAK))
    print("%9d  %s" % (per, "  ".join(cells)), flush=True)
