"""-m gpu: every libb200gan entry point against its CPU restatement (tests/abi_emul.py) or the torch op it replaces,
called through the C ABI exactly as the product does.  Tolerances: fp32 paths 1e-5 relative (summation order only),
tcgen05 bf16-operand paths 2e-2 relative (stated bf16 bound), index work bit-exact."""
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from abi_emul import EmulKernels  # noqa: E402
from b200gan import _lib, ops  # noqa: E402
from oracle import gan_oracle as O  # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
E = EmulKernels()


@pytest.fixture(scope="module")
def K():
    assert torch.cuda.is_available()
    return _lib.Kernels()


def cu(*ts):
    return [t.cuda() if isinstance(t, torch.Tensor) else t for t in ts]


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def close(a, b, tol, what=""):
    e = rel(a, b)
    assert e <= tol, "%s: rel err %.3e > %.1e" % (what, e, tol)


# ---------------------------------------------------------------------------------------------------------
# crops
# ---------------------------------------------------------------------------------------------------------
def _crop_inputs(seed=0, N=3, C=3, H=64, W=64, B=11):
    g = torch.Generator().manual_seed(seed)
    feats = torch.randn(N, C, H, W, generator=g)
    xy0 = torch.rand(B, 2, generator=g) * 0.6
    boxes = torch.cat([xy0, (xy0 + torch.rand(B, 2, generator=g) * 0.4 + 0.05).clamp(max=1.0)], 1)
    boxes[0] = torch.tensor([0.0, 0.0, 1.0, 1.0])
    boxes[1] = torch.tensor([0.5, 0.5, 0.5, 0.5])
    b2i = torch.sort(torch.randint(0, N, (B,), generator=g))[0]
    return feats, boxes, b2i


@pytest.mark.parametrize("HH,WW,H", [(32, 32, 64), (64, 64, 128), (16, 24, 48)])
def test_crop_taps_bit_exact(K, HH, WW, H):
    feats, boxes, b2i = _crop_inputs(H=H, W=H)
    wx, wy = ops.crop_weights(WW, "cuda"), ops.crop_weights(HH, "cuda")
    ix0, iy0, fx, fy = K.crop_taps(boxes.cuda(), wx, wy, H, H, HH, WW)
    rx0, ry0, rfx, rfy = O.crop_taps(boxes, H, H, HH, WW)
    assert torch.equal(ix0.cpu(), rx0) and torch.equal(iy0.cpu(), ry0), "floor indices must be bit-exact"
    assert torch.equal(fx.cpu(), rfx) and torch.equal(fy.cpu(), rfy), "fractional weights must be bit-exact"


def test_crop_fwd_bwd_vs_reference_golden(K):
    g = torch.load(os.path.join(GOLD, "crop.pt"))
    feats = g["feats"].cuda().requires_grad_(True)
    crops = ops.crop_bbox_batch(feats, g["boxes"].cuda(), g["b2f"], 32)
    assert float((crops.cpu() - g["crops"]).abs().max()) < 2e-6
    (crops * g["w"].cuda()).sum().backward()
    assert float((feats.grad.cpu() - g["dfeats"]).abs().max()) < 2e-5
    cu_ = ops.crop_bbox_batch(g["feats"].cuda(), g["boxes"].cuda(), g["b2f_u"], 16, 24)     # unsorted mapping
    assert float((cu_.cpu() - g["crops_u"]).abs().max()) < 2e-6


def test_crop_edge_cases(K):
    feats, boxes, b2i = _crop_inputs(N=4, B=5)
    b2i = torch.tensor([0, 0, 0, 3, 3])            # images 1 and 2 own no box
    x = feats.cuda().requires_grad_(True)
    y = ops.crop_bbox_batch(x, boxes[:5].cuda(), b2i, 32)
    ref_in = feats.clone().requires_grad_(True)
    ref = O.crop_bbox_batch(ref_in, boxes[:5], b2i, 32)
    close(y, ref, 1e-6, "crop fwd")
    y.sum().backward()
    ref.sum().backward()
    assert float((x.grad.cpu() - ref_in.grad).abs().max()) < 1e-4
    assert float(x.grad[1].abs().max()) == 0.0 and float(x.grad[2].abs().max()) == 0.0


# ---------------------------------------------------------------------------------------------------------
# convolutions (forward / dgrad / wgrad through the ops layer, i.e. descriptors + packing + kernels)
# ---------------------------------------------------------------------------------------------------------
CONV_CASES = [
    # Cx, Cy, k, s, p, H, N, x_layout, out_layout, bias
    (3, 64, 7, 1, 3, 32, 3, "nchw", "cl", False),
    (3, 64, 3, 1, 1, 16, 2, "nchw", "cl", True),
    (3, 64, 1, 1, 0, 16, 5, "nchw", "cl", False),       # OptimizedBlock shortcut on the pooled image
    (3, 128, 3, 2, 1, 17, 3, "nchw", "cl", False),      # strided, odd extent (im2col-packed path)
    (64, 128, 4, 2, 1, 32, 3, "cl", "cl", False),
    (64, 128, 4, 2, 1, 33, 2, "cl", "cl", False),       # odd extent (LayoutEncoder 66 -> 33 -> 16)
    (128, 64, 3, 1, 1, 8, 5, "cl", "cl", True),
    (64, 64, 1, 1, 0, 16, 2, "cl", "cl", True),
    (64, 3, 7, 1, 3, 16, 2, "cl", "nchw", True),
    (192, 256, 3, 1, 1, 8, 2, "cl", "cl", False),
    (128, 128, 5, 1, 2, 16, 1, "cl", "cl", False),
    (256, 512, 4, 2, 1, 8, 4, "cl", "cl", False),
]


def _conv_ref(x_nchw, w, b, s, p, relu, scale):
    y = F.conv2d(x_nchw, w, None, stride=s, padding=p)
    if scale is not None:
        y = y * scale
    if b is not None:
        y = y + b.view(1, -1, 1, 1)
    return F.relu(y) if relu else y


def _to_layout(t_nchw, layout):
    return t_nchw.contiguous() if layout == "nchw" else t_nchw.permute(0, 2, 3, 1).contiguous()


def _from_layout(t, layout):
    return t if layout == "nchw" else t.permute(0, 3, 1, 2)


def _bf(t):
    return t.bfloat16().float()


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_fwd_bwd(K, case, precision):
    """fp32: against F.conv2d autograd (with bias + fused ReLU).  bf16: the tcgen05 kernels round both operands to bf16
    and accumulate in fp32, so against the SAME computation on bf16-rounded operands they must agree to summation order
    (1e-4); layers the tensor path does not take (3-channel side) fall back to fp32 kernels and are held to 2e-2."""
    Cx, Cy, k, s, p, H, N, xl, ol, has_b = case
    g = torch.Generator().manual_seed(Cx * 7 + Cy + k)
    x = torch.randn(N, Cx, H, H, generator=g)
    w = torch.randn(Cy, Cx, k, k, generator=g) / (Cx * k * k) ** 0.5
    b = torch.randn(Cy, generator=g) if has_b else None
    relu = precision == "fp32"
    ops.set_precision(precision)
    try:
        geom = ops.ConvGeom(Cx, Cy, k, k, s, p)
        packs = ops.WeightPacks()
        xd = _to_layout(x, xl).cuda().requires_grad_(True)
        wd = w.cuda().requires_grad_(True)
        bd = b.cuda().requires_grad_(True) if has_b else None
        y = ops.conv2d(xd, wd, bd, geom, packs, xl, ol, relu=relu)
        if precision == "fp32":
            tf = td = tw = 1e-5
            xr, wr = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
            br = b.clone().requires_grad_(True) if has_b else None
            yr = _conv_ref(xr, wr, br, s, p, relu, None)
            gy = torch.randn(yr.shape, generator=g)
            yr.backward(gy)
            dxr, dwr, dbr = xr.grad, wr.grad, (br.grad if has_b else None)
        else:
            packed = ops._packed_ok(geom) and ol == "cl"     # few-channel input: im2col-packed tcgen05 path, bf16 output
            tf = 1e-4 if ops._tc_fwd_ok(geom, xl) else (4e-3 if packed else 2e-2)
            td = 1e-4 if ops._tc_dgrad_ok(geom, ol) else 2e-2
            tw = 1e-4 if (ops._tc_wgrad_ok(geom, xl, ol) or packed) else 2e-2
            xq, wq = (_bf(x), _bf(w)) if (ops._tc_fwd_ok(geom, xl) or packed) else (x, w)
            yr = _conv_ref(xq, wq, b, s, p, False, None)
            gy = torch.randn(yr.shape, generator=g)
            dxr = torch.nn.grad.conv2d_input(x.shape, _bf(w), _bf(gy), stride=s, padding=p)
            dwr = torch.nn.grad.conv2d_weight(_bf(x), w.shape, _bf(gy), stride=s, padding=p)
            dbr = gy.sum(dim=(0, 2, 3)) if has_b else None
        close(_from_layout(y, ol), yr, tf, "fwd")
        y.backward(_to_layout(gy, ol).cuda())
        close(_from_layout(xd.grad, xl), dxr, td, "dgrad")
        close(wd.grad, dwr, tw, "wgrad")
        if has_b:      # a bf16-stored output (NCHW input starting a bf16 network) receives its gradient rounded to bf16
            close(bd.grad, dbr, 1e-5 if y.dtype == torch.float32 else 4e-3, "bias grad")
    finally:
        ops.set_precision("fp32")


@pytest.mark.parametrize("case", [(3, 7, 1, 3, 32, 5, "nchw"), (3, 3, 1, 1, 16, 2, "nchw"), (3, 1, 1, 0, 9, 4, "nchw"),
                                  (3, 3, 2, 1, 17, 3, "nchw"), (3, 4, 2, 1, 8, 3, "cl")])
def test_im2col_pack_bit_exact(K, case):
    """the packed bf16 im2col matrix of a few-channel input equals the CPU restatement (unfold) bit for bit"""
    Cx, k, s, p, H, N, layout = case
    g = torch.Generator().manual_seed(Cx + k + H)
    x = torch.randn(N, Cx, H, H + 3, generator=g)
    geom = ops.ConvGeom(Cx, 64, k, k, s, p)
    Hy, Wy = geom.out_hw(H, H + 3)
    Kp = (Cx * k * k + 63) // 64 * 64
    xd = _to_layout(x, layout)
    strides = ops.nchw_strides(Cx, H, H + 3) if layout == "nchw" else ops.cl_strides(H, H + 3, Cx)
    got = K.im2col_pack(xd.cuda(), strides, N, H, H + 3, Cx, k, k, s, p, Hy, Wy, Kp)
    want = E.im2col_pack(xd, strides, N, H, H + 3, Cx, k, k, s, p, Hy, Wy, Kp)
    assert torch.equal(got.cpu().view(torch.int16), want.view(torch.int16))
    if s == 1 and layout == "nchw":
        # output rows made of whole 32-pixel strips: the shared-memory patch kernel (the shapes of the training step)
        for (Hh, Ww) in ((5, 32), (32, 64), (7, 96)):
            if 2 * p - k + 1 != 0:
                continue
            xs = torch.randn(N, Cx, Hh, Ww, generator=g)
            st = ops.nchw_strides(Cx, Hh, Ww)
            for flip in (False, True):
                got = K.im2col_pack(xs.cuda(), st, N, Hh, Ww, Cx, k, k, 1, p, Hh, Ww, Kp, flip=flip)
                want = E.im2col_pack(xs, st, N, Hh, Ww, Cx, k, k, 1, p, Hh, Ww, Kp, flip=flip)
                assert torch.equal(got.cpu().view(torch.int16), want.view(torch.int16)), (Hh, Ww, flip)
    if s == 1:      # transposed window walk (the im2col matrix of an output gradient), rows over the conv INPUT grid
        Hi, Wi = H + 2 * p - k + 1, H + 3 + 2 * p - k + 1
        dyv = torch.randn(N, Cx, Hi, Wi, generator=g)
        dyd = _to_layout(dyv, layout)
        st = ops.nchw_strides(Cx, Hi, Wi) if layout == "nchw" else ops.cl_strides(Hi, Wi, Cx)
        got = K.im2col_pack(dyd.cuda(), st, N, Hi, Wi, Cx, k, k, 1, p, H, H + 3, Kp, flip=True)
        want = E.im2col_pack(dyd, st, N, Hi, Wi, Cx, k, k, 1, p, H, H + 3, Kp, flip=True)
        assert torch.equal(got.cpu().view(torch.int16), want.view(torch.int16))


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("Cin,Cout,H", [(256, 256, 8), (256, 128, 16), (128, 64, 8)])
def test_conv_transpose(K, Cin, Cout, H, precision):
    tol = 1e-5 if precision == "fp32" else 1e-4
    g = torch.Generator().manual_seed(Cin + Cout)
    x = torch.randn(2, Cin, H, H, generator=g)
    w = torch.randn(Cin, Cout, 4, 4, generator=g) / (Cin * 4) ** 0.5
    q = _bf if precision == "bf16" else (lambda t: t)
    ops.set_precision(precision)
    try:
        geom = ops.ConvGeom(Cout, Cin, 4, 4, 2, 1)
        xd = _to_layout(x, "cl").cuda().requires_grad_(True)
        wd = w.cuda().requires_grad_(True)
        y = ops.conv_transpose2d(xd, wd, geom, ops.WeightPacks(), (2 * H, 2 * H))
        yr = F.conv_transpose2d(q(x), q(w), None, stride=2, padding=1)
        close(_from_layout(y, "cl"), yr, tol, "convT fwd")
        gy = torch.randn(yr.shape, generator=g)
        y.backward(_to_layout(gy, "cl").cuda())
        dxr = F.conv2d(q(gy), q(w), None, stride=2, padding=1)
        dwr = torch.nn.grad.conv2d_weight(q(gy), (Cin, Cout, 4, 4), q(x), stride=2, padding=1)
        close(_from_layout(xd.grad, "cl"), dxr, tol, "convT dgrad")
        close(wd.grad, dwr, tol, "convT wgrad")
    finally:
        ops.set_precision("fp32")


def test_linear_and_spectral_norm(K):
    g = torch.Generator().manual_seed(3)
    x = torch.randn(37, 170, generator=g)
    w = torch.randn(128, 170, generator=g) * 0.1
    b = torch.randn(128, generator=g)
    u = F.normalize(torch.randn(128, generator=g), dim=0)
    v = F.normalize(torch.randn(170, generator=g), dim=0)
    xd, wd, bd = x.cuda().requires_grad_(True), w.cuda().requires_grad_(True), b.cuda().requires_grad_(True)
    ud, vd = u.cuda(), v.cuda()
    y = ops.linear(xd, wd, bd, ops.WeightPacks(), sn=ops.sn_iterate(wd, ud, vd, 1, True))
    st = {"l.weight_orig": w.clone().requires_grad_(True), "l.weight_u": u.clone(), "l.weight_v": v.clone(), "l.bias": b.clone().requires_grad_(True)}
    xr = x.clone().requires_grad_(True)
    yr = F.linear(xr, O.sn_weight(st, "l", True), st["l.bias"])
    close(y, yr, 1e-5, "sn linear fwd")
    close(ud, st["l.weight_u"], 1e-6, "u")
    close(vd, st["l.weight_v"], 1e-6, "v")
    gy = torch.randn(yr.shape, generator=g)
    y.backward(gy.cuda())
    yr.backward(gy)
    close(xd.grad, xr.grad, 1e-5, "sn dgrad")
    close(wd.grad, st["l.weight_orig"].grad, 2e-5, "sn wgrad")
    close(bd.grad, st["l.bias"].grad, 1e-5, "bias")


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_batched_spectral_norm_conv(K, precision):
    """three calls of a spectral-normalised conv batched along dim 0 == three sequential calls of the reference hook:
    per-call sigma in the epilogue / dgrad, per-call gradient through sigma, u / v advanced three times."""
    g = torch.Generator().manual_seed(5)
    groups, n, Cin, Cout, H = 3, 2, 64, 128, 8
    x = torch.randn(groups * n, Cin, H, H, generator=g)
    w = torch.randn(Cout, Cin, 3, 3, generator=g) * 0.05
    b = torch.randn(Cout, generator=g)
    u = F.normalize(torch.randn(Cout, generator=g), dim=0)
    v = F.normalize(torch.randn(Cin * 9, generator=g), dim=0)
    gy = torch.randn(groups * n, Cout, H, H, generator=g)
    tol = 2e-2 if precision == "bf16" else 2e-5          # bf16: tcgen05 operands rounded to bf16, reference kept in fp32
    relu = precision == "fp32"                           # (a fused ReLU would flip near-zero mask bits under bf16 rounding)
    ops.set_precision(precision)
    try:
        xd = _to_layout(x, "cl").cuda().requires_grad_(True)
        wd, bd, ud, vd = w.cuda().requires_grad_(True), b.cuda().requires_grad_(True), u.cuda(), v.cuda()
        geom = ops.ConvGeom(Cin, Cout, 3, 3, 1, 1)
        y = ops.conv2d(xd, wd, bd, geom, ops.WeightPacks(), "cl", "cl", relu=relu, sn=ops.sn_iterate(wd, ud, vd, groups, True))
        y.backward(_to_layout(gy, "cl").cuda())
    finally:
        ops.set_precision("fp32")
    st = {"l.weight_orig": w.clone().requires_grad_(True), "l.weight_u": u.clone(), "l.weight_v": v.clone()}
    bias = b.clone().requires_grad_(True)
    xr = x.clone().requires_grad_(True)
    ys = []
    for gi in range(groups):
        wn = O.sn_weight(st, "l", True)
        yg = F.conv2d(xr[gi * n:(gi + 1) * n], wn, bias, padding=1)
        ys.append(F.relu(yg) if relu else yg)
    yr = torch.cat(ys)
    yr.backward(gy)
    close(_from_layout(y, "cl"), yr, tol, "batched sn conv fwd")
    close(ud, st["l.weight_u"], 1e-5, "u after 3 iterations")
    close(vd, st["l.weight_v"], 1e-5, "v after 3 iterations")
    close(_from_layout(xd.grad, "cl"), xr.grad, tol, "batched sn dgrad")
    close(wd.grad, st["l.weight_orig"].grad, tol if precision == "bf16" else 5e-5, "batched sn wgrad")
    close(bd.grad, bias.grad, 1e-5, "bias")


def test_whole_network_power_iteration(K):
    """b200_sn_power_iter_multi (all layers of a discriminator, 3 iterations) == the per-layer kernels run 3 times:
    1/sigma per iteration, u / v histories and the final in-place u / v state, bit for bit (same reduction order)."""
    from models.discriminator import ObjectDiscriminator
    from b200gan import nn as bnn
    torch.manual_seed(0)
    net = bnn.add_sn(ObjectDiscriminator(n_class=179)).cuda()
    mods = [m for m in net.modules() if getattr(m, "_b200_sn", False)]
    assert len(mods) == 17
    ref = []
    for m in mods:
        u, v = m.weight_u.clone(), m.weight_v.clone()
        ref.append((ops.sn_iterate(m.weight_orig, u, v, 3, True), u, v))
    bnn.sn_prepare(net, 3)
    for m, (call, u, v) in zip(mods, ref):
        got = m.__dict__["_sn_staged"]
        assert torch.equal(got.inv, call.inv) and torch.equal(got.u_hist, call.u_hist) and torch.equal(got.v_hist, call.v_hist)
        assert torch.equal(m.weight_u, u) and torch.equal(m.weight_v, v)
    net.eval()                                    # eval mode: sigma from the current u, v without iterating
    before = [m.weight_u.clone() for m in mods]
    bnn.sn_prepare(net, 1)
    for m, u0 in zip(mods, before):
        assert torch.equal(m.weight_u, u0)
        W = m.weight_orig.reshape(m.weight_orig.shape[0], -1)
        sigma = torch.dot(m.weight_u, W @ m.weight_v)
        close(m.__dict__["_sn_staged"].inv, (1.0 / sigma).reshape(1), 1e-5, "eval sigma")


@pytest.mark.parametrize("case", [(64, 64, 3, 1, 1, 16, 3), (64, 128, 3, 1, 1, 8, 4), (128, 64, 1, 1, 0, 16, 4),
                                  (256, 512, 3, 1, 1, 4, 4)])
def test_grouped_sn_weight_gradient(K, case):
    """`groups` batched calls of a spectral-normalised conv: ONE weight-gradient GEMM with call-aligned splits +
    b200_sn_wgrad_finish must equal the per-call loop (wgrad + reduce + b200_sn_grad per call) up to summation order, and
    the emulation of the entry point."""
    Cx, Cy, k, s, p, H, groups = case
    n = 16                                                   # images per call
    g = torch.Generator().manual_seed(Cx + Cy + groups)
    geom = ops.ConvGeom(Cx, Cy, k, k, s, p)
    Hy = geom.out_hw(H, H)[0]
    x = torch.randn(groups * n, H, H, Cx, generator=g)
    dy = torch.randn(groups * n, Hy, Hy, Cy, generator=g)
    w = (torch.randn(Cy, Cx, k, k, generator=g) / (Cx * k * k) ** 0.5)
    u = F.normalize(torch.randn(Cy, generator=g), dim=0)
    v = F.normalize(torch.randn(Cx * k * k, generator=g), dim=0)
    ops.set_precision("bf16")
    try:
        outs = []
        for grouped in (True, False):
            prev, ops.GROUPED_SN_WGRAD = ops.GROUPED_SN_WGRAD, grouped
            try:
                wd = w.cuda().requires_grad_(True)
                sn = ops.sn_iterate(wd.detach(), u.cuda(), v.cuda(), groups, True)
                if grouped:
                    assert ops._sn_group_splits(geom, groups * n * Hy * Hy, groups) > 0
                y = ops.conv2d(x.cuda().bfloat16(), wd, None, geom, ops.WeightPacks(), sn=sn)
                y.backward(dy.cuda().bfloat16())
                outs.append(wd.grad.clone())
            finally:
                ops.GROUPED_SN_WGRAD = prev
        close(outs[0], outs[1], 2e-5, "grouped vs per-call spectral-norm weight gradient")
    finally:
        ops.set_precision("fp32")


@pytest.mark.parametrize("h,w", [(64, 27), (1, 1024), (179, 1024), (1024, 9216)])
def test_sn_power_iteration(K, h, w):
    g = torch.Generator().manual_seed(h + w)
    W = torch.randn(h, w, generator=g)
    u = F.normalize(torch.randn(h, generator=g), dim=0)
    v = F.normalize(torch.randn(w, generator=g), dim=0)
    ud, vd = u.cuda(), v.cuda()
    for do_iter in (1, 1, 0):
        s_gpu = K.sn_power_iter(W.cuda(), h, w, ud, vd, do_iter, 1e-12)
        s_cpu = E.sn_power_iter(W, h, w, u, v, do_iter, 1e-12)
        close(s_gpu, s_cpu, 2e-5, "sigma")
        close(ud, u, 2e-5, "u")
        close(vd, v, 2e-5, "v")


# ---------------------------------------------------------------------------------------------------------
# normalisation
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("groups", [1, 3])
@pytest.mark.parametrize("mode", [0, 1, 2, 3])
@pytest.mark.parametrize("relu", [False, True])
def test_norm_fwd_bwd(K, mode, relu, groups):
    """groups > 1: independent calls batched along the rows keep separate statistics (and sequential running-stat
    updates); the emulation restates exactly that."""
    g = torch.Generator().manual_seed(10 + mode)
    O_, hw, C, ncls = 6 * groups, 20, 64, 9
    rows = O_ * hw
    x = torch.randn(rows, C, generator=g) * 2 + 0.5
    idx = torch.randint(0, ncls, (O_,), generator=g).to(torch.int32)
    if mode == 1:
        gamma, beta = torch.randn(C, generator=g), torch.randn(C, generator=g)
    elif mode == 2:
        gamma, beta = torch.randn(ncls, 2 * C, generator=g), None
    elif mode == 3:
        gamma, beta = torch.randn(rows, 2 * C, generator=g), None
    else:
        gamma = beta = None
    rm, rv = torch.zeros(C), torch.ones(C)
    rmd, rvd = rm.cuda(), rv.cuda()
    x = x + torch.arange(groups).repeat_interleave(rows // groups)[:, None].float()      # distinct group statistics
    mean_d, var_d = K.bn_stats(x.cuda(), rmd, rvd, 0.1, groups)
    mean_c, var_c = E.bn_stats(x, rm, rv, 0.1, groups)
    if groups > 1:
        sep = [E.bn_stats(x[i * rows // groups:(i + 1) * rows // groups], None, None, 0.1)[0] for i in range(groups)]
        close(mean_c, torch.cat(sep), 1e-6, "grouped stats == separate calls")
    close(mean_d, mean_c, 1e-6, "mean")
    close(var_d, var_c, 1e-6, "var")
    close(rmd, rm, 1e-6, "running_mean")
    close(rvd, rv, 1e-6, "running_var")
    res = torch.randn(rows, C, generator=g) if (mode == 1 and not relu) else None
    a = cu(x, mean_c, var_c)
    yd = K.norm_fwd(a[0], a[1], a[2], 1e-5, mode, *cu(gamma, beta, idx if mode == 2 else None), hw, *cu(res), relu, groups)
    yc = E.norm_fwd(x, mean_c, var_c, 1e-5, mode, gamma, beta, idx, hw, res, relu, groups)
    close(yd, yc, 1e-6, "norm fwd")
    dy = torch.randn(rows, C, generator=g)
    outs_d = K.norm_bwd(*cu(dy, x, yc, mean_c, var_c), 1e-5, mode, *cu(gamma, idx if mode == 2 else None), hw, relu, ncls, groups)
    outs_c = E.norm_bwd(dy, x, yc, mean_c, var_c, 1e-5, mode, gamma, idx, hw, relu, ncls, groups)
    for name, d, c in zip(("dx", "dgamma", "dbeta", "dtable", "dgb"), outs_d, outs_c):
        assert (d is None) == (c is None), name
        if d is not None:
            close(d, c, 2e-5, name)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_cbn_recomputed_relu_mask(K, dtype):
    """relu = 2: conditional batch norm's backward recomputes the ReLU mask from x (gamma * xhat + beta > 0) instead of
    reading y; it must equal the y-based mask path EXACTLY (same fp32 expression as the forward kernel)."""
    g = torch.Generator().manual_seed(3)
    groups, O_, hw, C, ncls = 3, 12, 36, 128, 9
    rows = O_ * hw
    x = (torch.randn(rows, C, generator=g) * 2 + 0.5).to(dtype)
    idx = torch.randint(0, ncls, (O_,), generator=g).to(torch.int32)
    table = torch.randn(ncls, 2 * C, generator=g)
    dy = torch.randn(rows, C, generator=g).to(dtype)
    xd, dyd, td, idd = cu(x, dy, table, idx)
    mean, var = K.bn_stats(xd, None, None, 0.1, groups)
    y = K.norm_fwd(xd, mean, var, 1e-5, 2, td, None, idd, hw, None, True, groups)
    a = K.norm_bwd(dyd, xd, y, mean, var, 1e-5, 2, td, idd, hw, 1, ncls, groups)
    b = K.norm_bwd(dyd, xd, None, mean, var, 1e-5, 2, td, idd, hw, 2, ncls, groups)
    assert torch.equal(a[0], b[0]) and torch.equal(a[3], b[3]), "recomputed mask differs from the saved-output mask"
    c = E.norm_bwd(dy, x, None, mean.cpu(), var.cpu(), 1e-5, 2, table, idx, hw, 2, ncls, groups)
    close(b[0], c[0], 2e-5 if dtype == torch.float32 else 1e-2, "dx")
    close(b[3], c[3], 2e-5 if dtype == torch.float32 else 1e-2, "dtable")


def test_bn_stats_shapes(K):
    for rows, C in ((1, 64), (7, 128), (100000, 64), (4096, 1024)):
        x = torch.randn(rows, C) + 3.0
        m, v = K.bn_stats(x.cuda(), None, None, 0.1)
        close(m, x.double().mean(0), 1e-6, "mean %d" % rows)
        if rows > 1:
            assert float((v.cpu() - x.double().var(0, unbiased=False).float()).abs().max()) < 1e-5


# ---------------------------------------------------------------------------------------------------------
# elementwise / pooling / layout
# ---------------------------------------------------------------------------------------------------------
def test_elementwise_family(K):
    g = torch.Generator().manual_seed(1)
    x = torch.randn(5, 8, 8, 64, generator=g)
    y = torch.randn(5, 8, 8, 64, generator=g)
    close(K.relu_fwd(x.cuda()), F.relu(x), 0, "relu")
    close(K.relu_bwd(y.cuda(), F.relu(x).cuda()), E.relu_bwd(y, F.relu(x)), 0, "relu bwd")
    close(K.add(x.cuda(), y.cuda()), x + y, 0, "add")
    odd = torch.randn(1003, generator=g)
    close(K.relu_fwd(odd.cuda()), F.relu(odd), 0, "relu tail")
    for f, sc in ((2, 0.25), (8, 1.0), (4, 1.0 / 16)):
        close(K.pool_fwd(x.cuda(), 5, 8, 8, 64, f, sc), E.pool_fwd(x, 5, 8, 8, 64, f, sc), 1e-6, "pool")
        close(K.unpool_fwd(x.cuda(), 5, 8, 8, 64, f, sc), E.unpool_fwd(x, 5, 8, 8, 64, f, sc), 1e-7, "unpool")
    a, b = torch.randn(6 * 4, 64, generator=g), torch.randn(6, 128, generator=g)
    close(K.concat_fwd(a.cuda(), 64, 1, b.cuda(), 128, 4, 24), E.concat_fwd(a, 64, 1, b, 128, 4, 24), 0, "concat")
    d = torch.randn(24, 192, generator=g)
    da, db = K.concat_bwd(d.cuda(), 64, 1, 128, 4, 24)
    ra, rb = E.concat_bwd(d, 64, 1, 128, 4, 24)
    close(da, ra, 1e-6, "concat bwd a")
    close(db, rb, 1e-6, "concat bwd b")
    table = torch.randn(179, 128, generator=g)
    idx = torch.randint(0, 179, (40,), generator=g).to(torch.int32)
    close(K.gather_rows(table.cuda(), idx.cuda()), table[idx.long()], 0, "gather")
    dd = torch.randn(40, 128, generator=g)
    close(K.scatter_rows(dd.cuda(), idx.cuda(), 179), E.scatter_rows(dd, idx, 179), 1e-6, "scatter")
    src = torch.tensor([3, -1, 0, 2, 2], dtype=torch.int32)
    xr = torch.randn(4, 64, generator=g)
    close(K.permute_rows(xr.cuda(), src.cuda(), 64), E.permute_rows(xr, src, 64), 0, "permute")
    v = torch.randn(7, 64, generator=g)
    mask = (torch.rand(7, 1, 16, 16, generator=g) > 0.5).float()
    close(K.mask_outer_fwd(v.cuda(), mask.cuda(), 7, 16, 16, 64), E.mask_outer_fwd(v, mask, 7, 16, 16, 64), 0, "mask outer")
    do = torch.randn(7, 18, 18, 64, generator=g)
    close(K.mask_outer_bwd(do.cuda(), mask.cuda(), 7, 16, 16, 64), E.mask_outer_bwd(do, mask, 7, 16, 16, 64), 1e-6, "mask outer bwd")
    mu, lv, eps = [torch.randn(9, 64, generator=g) for _ in range(3)]
    close(K.reparam_fwd(*cu(mu, lv, eps)), E.reparam_fwd(mu, lv, eps), 1e-6, "reparam")
    for dgpu, dcpu in zip(K.reparam_bwd(*cu(mu, lv, eps)), E.reparam_bwd(mu, lv, eps)):
        close(dgpu, dcpu, 1e-6, "reparam bwd")
    t = torch.randn(3, 50, 7, generator=g)
    close(K.transpose(t.cuda(), 3, 50, 7), t.transpose(1, 2), 0, "transpose")
    close(K.colsum(d.cuda()), d.sum(0), 1e-6, "colsum")


def test_bf16_activation_storage(K):
    """bf16 mode stores channel-last activations as bf16: the same kernels read bf16, compute in fp32 and round the result
    once.  Against the emulation (which does exactly that) the only freedom is the last bf16 bit of a few elements."""
    g = torch.Generator().manual_seed(21)
    bf = lambda t: t.to(torch.bfloat16)
    tol = 1e-3
    x, y = bf(torch.randn(5, 8, 8, 64, generator=g)), bf(torch.randn(5, 8, 8, 64, generator=g))
    assert K.relu_fwd(x.cuda()).dtype == torch.bfloat16
    close(K.relu_fwd(x.cuda()), F.relu(x), 0, "relu")
    close(K.relu_bwd(y.cuda(), F.relu(x).cuda()), E.relu_bwd(y, F.relu(x)), 0, "relu bwd")
    close(K.add(x.cuda(), y.cuda()), E.add(x, y), tol, "add")
    odd = bf(torch.randn(1003, generator=g))
    close(K.relu_fwd(odd.cuda()), F.relu(odd), 0, "relu tail")
    for f, sc in ((2, 0.25), (8, 1.0)):
        close(K.pool_fwd(x.cuda(), 5, 8, 8, 64, f, sc), E.pool_fwd(x, 5, 8, 8, 64, f, sc), tol, "pool")
        close(K.unpool_fwd(x.cuda(), 5, 8, 8, 64, f, sc), E.unpool_fwd(x, 5, 8, 8, 64, f, sc), tol, "unpool")
    a, b = bf(torch.randn(24, 64, generator=g)), bf(torch.randn(6, 128, generator=g))
    close(K.concat_fwd(a.cuda(), 64, 1, b.cuda(), 128, 4, 24), E.concat_fwd(a, 64, 1, b, 128, 4, 24), 0, "concat")
    d = bf(torch.randn(24, 192, generator=g))
    for dg, dc in zip(K.concat_bwd(d.cuda(), 64, 1, 128, 4, 24), E.concat_bwd(d, 64, 1, 128, 4, 24)):
        close(dg, dc, tol, "concat bwd")
    src = torch.tensor([3, -1, 0, 2, 2], dtype=torch.int32)
    xr = bf(torch.randn(4, 64, generator=g))
    close(K.permute_rows(xr.cuda(), src.cuda(), 64), E.permute_rows(xr, src, 64), 0, "permute")
    v = torch.randn(7, 64, generator=g)
    mask = (torch.rand(7, 1, 16, 16, generator=g) > 0.5).float()
    mo = K.mask_outer_fwd(v.cuda(), mask.cuda(), 7, 16, 16, 64, torch.bfloat16)
    assert mo.dtype == torch.bfloat16
    close(mo, E.mask_outer_fwd(v, mask, 7, 16, 16, 64, torch.bfloat16), 0, "mask outer")
    do = bf(torch.randn(7, 18, 18, 64, generator=g))
    close(K.mask_outer_bwd(do.cuda(), mask.cuda(), 7, 16, 16, 64), E.mask_outer_bwd(do, mask, 7, 16, 16, 64), 1e-5, "mask outer bwd")
    close(K.colsum(d.cuda()), d.float().sum(0), 1e-6, "colsum")
    # ConvLSTM gates: bf16 pre-activations / hidden state, fp32 cell state and saved gates
    rows, hid = 3 * 64, 64
    px, ph = bf(torch.randn(rows, 4 * hid, generator=g)), bf(torch.randn(rows, 4 * hid, generator=g))
    cp = torch.randn(rows, hid, generator=g)
    gd, cd, hd = K.lstm_gates_fwd(*cu(px, ph, cp), rows, hid)
    gc, cc, hc = E.lstm_gates_fwd(px, ph, cp, rows, hid)
    assert hd.dtype == torch.bfloat16 and gd.dtype == torch.float32 and cd.dtype == torch.float32
    close(gd, gc, 1e-5, "gates"); close(cd, cc, 1e-5, "c"); close(hd, hc, tol, "h")
    dh, dcn = bf(torch.randn(rows, hid, generator=g)), torch.randn(rows, hid, generator=g)
    dpd, dcd = K.lstm_gates_bwd(*cu(dh, dcn, gc, cp, cc), rows, hid)
    dpc, dcc = E.lstm_gates_bwd(dh, dcn, gc, cp, cc, rows, hid)
    close(dpd, dpc, tol, "dpre"); close(dcd, dcc, 1e-5, "dc_prev")
    # normalisation family (grouped statistics, all four modes)
    O_, hw, C, ncls, groups = 6, 20, 64, 9, 3
    rows = O_ * hw
    xn = bf(torch.randn(rows, C, generator=g) * 2 + 0.5)
    idx = torch.randint(0, ncls, (O_,), generator=g).to(torch.int32)
    mean_d, var_d = K.bn_stats(xn.cuda(), None, None, 0.1, groups)
    mean_c, var_c = E.bn_stats(xn, None, None, 0.1, groups)
    close(mean_d, mean_c, 1e-6, "mean"); close(var_d, var_c, 1e-6, "var")
    for mode in (0, 1, 2, 3):
        if mode == 1:
            gamma, beta = torch.randn(C, generator=g), torch.randn(C, generator=g)
        elif mode == 2:
            gamma, beta = torch.randn(ncls, 2 * C, generator=g), None
        elif mode == 3:
            gamma, beta = bf(torch.randn(rows, 2 * C, generator=g) * 0.3), None
        else:
            gamma = beta = None
        res = bf(torch.randn(rows, C, generator=g)) if mode == 1 else None
        relu = mode != 1
        yd = K.norm_fwd(*cu(xn, mean_c, var_c), 1e-5, mode, *cu(gamma, beta, idx if mode == 2 else None), hw, *cu(res), relu, groups)
        yc = E.norm_fwd(xn, mean_c, var_c, 1e-5, mode, gamma, beta, idx, hw, res, relu, groups)
        assert yd.dtype == torch.bfloat16
        close(yd, yc, tol, "norm fwd mode %d" % mode)
        dy = bf(torch.randn(rows, C, generator=g))
        outs_d = K.norm_bwd(*cu(dy, xn, yc, mean_c, var_c), 1e-5, mode, *cu(gamma, idx if mode == 2 else None), hw, relu, ncls, groups)
        outs_c = E.norm_bwd(dy, xn, yc, mean_c, var_c, 1e-5, mode, gamma, idx, hw, relu, ncls, groups)
        for name, dd, c in zip(("dx", "dgamma", "dbeta", "dtable", "dgb"), outs_d, outs_c):
            assert (dd is None) == (c is None), name
            if dd is not None:
                close(dd, c, tol if name in ("dx", "dgb") else 2e-5, "%s mode %d" % (name, mode))


@pytest.mark.parametrize("case", [(64, 128, 3, 1, 1, 8, 2, "cl", "cl"), (3, 64, 7, 1, 3, 16, 2, "nchw", "cl"),
                                  (64, 3, 7, 1, 3, 16, 2, "cl", "nchw"), (128, 64, 4, 2, 1, 16, 2, "cl", "cl")])
def test_conv_bf16_activation_storage(K, case):
    """bf16 mode end to end for one conv: channel-last activations and their gradients are bf16 tensors (NCHW boundary
    tensors fp32), on the tcgen05 kernels and on the fp32 CUDA-core kernels of the 3-channel layers alike."""
    Cx, Cy, k, s, p, H, N, xl, ol = case
    g = torch.Generator().manual_seed(Cx + 5 * Cy)
    x = torch.randn(N, Cx, H, H, generator=g)
    w = torch.randn(Cy, Cx, k, k, generator=g) / (Cx * k * k) ** 0.5
    b = torch.randn(Cy, generator=g)
    ops.set_precision("bf16")
    try:
        geom = ops.ConvGeom(Cx, Cy, k, k, s, p)
        xin = _to_layout(x, xl)
        xin = xin.to(torch.bfloat16) if xl == "cl" else xin
        xd = xin.cuda().requires_grad_(True)
        wd, bd = w.cuda().requires_grad_(True), b.cuda().requires_grad_(True)
        y = ops.conv2d(xd, wd, bd, geom, ops.WeightPacks(), xl, ol)
        assert y.dtype == (torch.bfloat16 if ol == "cl" else torch.float32)
        xr = xin.float().requires_grad_(True) if xl == "nchw" else _from_layout(xin.float(), "cl").contiguous().requires_grad_(True)
        wr, br = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
        yr = F.conv2d(xr, wr, br, stride=s, padding=p)
        close(_from_layout(y.float(), ol), yr, 1e-2, "fwd")
        gy = torch.randn(yr.shape, generator=g)
        gyl = _to_layout(gy, ol)
        gyl = gyl.to(torch.bfloat16) if ol == "cl" else gyl
        y.backward(gyl.cuda())
        yr.backward(_from_layout(gyl.float(), ol))
        assert xd.grad.dtype == xd.dtype
        close(_from_layout(xd.grad.float(), xl), xr.grad, 1e-2, "dgrad")
        close(wd.grad, wr.grad, 1e-2, "wgrad")
        close(bd.grad, br.grad, 1e-5, "bias grad")
    finally:
        ops.set_precision("fp32")


def test_lstm_gates(K):
    g = torch.Generator().manual_seed(2)
    rows, hid = 3 * 64, 64
    px, ph, cp = torch.randn(rows, 4 * hid, generator=g), torch.randn(rows, 4 * hid, generator=g), torch.randn(rows, hid, generator=g)
    for with_prev in (True, False):
        a = (px, ph, cp) if with_prev else (px, None, None)
        gd, cd, hd = K.lstm_gates_fwd(*cu(*a), rows, hid)
        gc, cc, hc = E.lstm_gates_fwd(*a, rows, hid)
        close(gd, gc, 1e-6, "gates")
        close(cd, cc, 1e-6, "c")
        close(hd, hc, 1e-6, "h")
        dh, dcn = torch.randn(rows, hid, generator=g), torch.randn(rows, hid, generator=g)
        b = (dh, dcn if with_prev else None, gc, cp if with_prev else None, cc)
        dpd, dcd = K.lstm_gates_bwd(*cu(*b), rows, hid)
        dpc, dcc = E.lstm_gates_bwd(*b, rows, hid)
        close(dpd, dpc, 1e-5, "dpre")
        close(dcd, dcc, 1e-5, "dc_prev")


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_conv_lstm_vs_oracle(K, precision):
    tol = 2e-5 if precision == "fp32" else 3e-2
    g = torch.Generator().manual_seed(4)
    o2i = torch.tensor([0, 0, 0, 1, 2, 2, 2, 2, 3, 3])
    x = torch.randn(10, 512, 8, 8, generator=g) * 0.5
    sd = {}
    cin = 512
    layers = []
    for i, hid in enumerate((128, 64, 64)):
        sd["c.cell_list.%d.conv.weight" % i] = (torch.randn(4 * hid, cin + hid, 5, 5, generator=g) / ((cin + hid) * 25) ** 0.5).requires_grad_(True)
        sd["c.cell_list.%d.conv.bias" % i] = (torch.randn(4 * hid, generator=g) * 0.1).requires_grad_(True)
        layers.append(ops.ConvLSTMLayer(cin, hid, 5))
        cin = hid
    xr = x.clone().requires_grad_(True)
    yr = O.conv_lstm(sd, "c", xr, o2i)
    gy = torch.randn(yr.shape, generator=g)
    yr.backward(gy)
    ops.set_precision(precision)
    try:
        xd = x.permute(0, 2, 3, 1).contiguous().cuda().requires_grad_(True)
        params = []
        for i in range(3):
            params += [sd["c.cell_list.%d.conv.weight" % i].detach().cuda().requires_grad_(True),
                       sd["c.cell_list.%d.conv.bias" % i].detach().cuda().requires_grad_(True)]
        plan = ops.get_plan(o2i, None, "cuda")
        y = ops.conv_lstm(xd, plan, layers, params)
        close(y.permute(0, 3, 1, 2), yr, tol, "clstm fwd")
        y.backward(gy.permute(0, 2, 3, 1).contiguous().cuda())
        close(xd.grad.permute(0, 3, 1, 2), xr.grad, tol * 2, "clstm dx")
        for i in range(3):
            close(params[2 * i].grad, sd["c.cell_list.%d.conv.weight" % i].grad, tol * 2, "clstm dW%d" % i)
            close(params[2 * i + 1].grad, sd["c.cell_list.%d.conv.bias" % i].grad, tol * 2, "clstm db%d" % i)
    finally:
        ops.set_precision("fp32")


def test_tc_matches_fp32_at_scale(K):
    """size-independent property at BASELINE size (O=256 crops, 64->128 k4s2 at 32x32): the tcgen05 result equals the fp32
    CUDA-core result within the bf16 operand bound, and conv is linear: conv(a*x) == a*conv(x)."""
    g = torch.Generator().manual_seed(5)
    x = torch.randn(256, 32, 32, 64, generator=g).cuda()
    w = (torch.randn(128, 64, 4, 4, generator=g) / 32).cuda()
    geom = ops.ConvGeom(64, 128, 4, 4, 2, 1)
    y32 = ops.conv_forward(geom, ops.WeightPacks(), w, x, "cl", "cl")
    ops.set_precision("bf16")
    try:
        ytc = ops.conv_forward(geom, ops.WeightPacks(), w, x, "cl", "cl")
        ytc2 = ops.conv_forward(geom, ops.WeightPacks(), w, x * 2.0, "cl", "cl")
    finally:
        ops.set_precision("fp32")
    close(ytc, y32, 1e-2, "tc vs fp32")
    assert torch.equal(ytc2, ytc * 2.0), "scaling by a power of two must commute exactly with the bf16 GEMM"


IM2COL_CASES = [
    # (Cx, Cy, k, s, p, H, N, transposed)
    (64, 64, 3, 1, 1, 64, 40, False),     # 1280 tiles: several tiles per persistent CTA, both TMEM buffers in use
    (256, 512, 3, 1, 1, 16, 20, False),   # 4 N tiles per M tile
    (64, 3, 7, 1, 3, 32, 24, False),      # BN = 16 tile (image convolution 64 -> 3)
    (64, 128, 3, 1, 1, 32, 9, False),     # shifted-window shapes: D blocks at 32x32 / 64x64, SPADE, 5x5
    (128, 128, 3, 1, 1, 33, 3, False),
    (128, 64, 5, 1, 2, 20, 3, False),
    (64, 64, 3, 1, 1, 128, 2, False),
    (64, 64, 3, 1, 1, 64, 3, False),      # D blocks at 64x64
    (128, 128, 3, 1, 1, 16, 5, False),
    (64, 128, 4, 2, 1, 66, 3, False),     # LayoutEncoder c2 (odd geometry 66 -> 33)
    (128, 256, 4, 2, 1, 33, 3, False),    # 33 -> 16
    (128, 256, 5, 1, 2, 8, 7, False),     # ConvLSTM gate conv
    (192, 256, 3, 1, 1, 8, 2, False),
    (512, 512, 1, 1, 0, 4, 9, False),     # 1x1 shortcut
    (128, 256, 4, 2, 1, 8, 3, True),      # ConvTranspose phases (dgrad descriptors, negative tap steps)
]


@pytest.mark.parametrize("case", IM2COL_CASES)
def test_im2col_path_equals_cpasync_path(K, case):
    """The persistent TMA-im2col kernel, the one-tile-per-CTA TMA-im2col kernel and the cp.async gather kernel fetch the
    same bf16 tiles and issue the same MMAs in the same order, so forward, dgrad and the ConvTranspose phases must agree
    BIT-EXACTLY between the three."""
    Cx, Cy, k, s, p, H, N, transposed = case
    g = torch.Generator().manual_seed(Cx + Cy * 3 + k)
    geom = ops.ConvGeom(Cx, Cy, k, k, s, p)
    w = (torch.randn(Cy, Cx, k, k, generator=g) / (Cx * k * k) ** 0.5).cuda()
    Hy = geom.out_hw(H, H)[0]
    x = torch.randn(N, H, H, Cx, generator=g).cuda()
    dy = torch.randn(N, Hy, Hy, Cy, generator=g).cuda()
    ops.set_precision("bf16")
    outs = []
    try:
        for enable, persist, halo in ((True, True, False), (False, False, False), (True, False, False), (True, True, True)):
            prev = _lib.K.conv_tc_set_im2col(enable)
            prev_p = _lib.K.conv_tc_set_persistent(persist)
            prev_h = _lib.K.conv_tc_set_halo(halo)
            try:
                packs = ops.WeightPacks()
                if not transposed:
                    y = ops.conv_forward(geom, packs, w, x, "cl", "cl")
                    dx = ops.conv_dgrad(geom, packs, w, dy, "cl", (H, H), "cl")
                else:
                    y = ops.conv_dgrad(geom, packs, w, dy, "cl", (H, H), "cl")
                    dx = ops.conv_forward(geom, packs, w, x, "cl", "cl")
                dw = ops.conv_wgrad(geom, x, "cl", dy, "cl", torch.empty_like(w))
                outs.append((y.clone(), dx.clone(), dw.clone()))
            finally:
                _lib.K.conv_tc_set_im2col(prev)
                _lib.K.conv_tc_set_persistent(prev_p)
                _lib.K.conv_tc_set_halo(prev_h)
    finally:
        ops.set_precision("fp32")
    # the shifted-window kernel accumulates channel-slab-major instead of tap-major: bit-exact for one 64-channel slab,
    # summation-order differences otherwise
    if Cx == 64 and Cy == 64:
        assert torch.equal(outs[3][0], outs[0][0]) and torch.equal(outs[3][1], outs[0][1]), "shifted-window kernel differs"
    close(outs[3][0], outs[0][0], 2e-6, "shifted-window forward")
    close(outs[3][1], outs[0][1], 2e-6, "shifted-window dgrad")
    assert torch.equal(outs[0][0], outs[2][0]), "forward differs between the persistent and the per-tile kernel"
    assert torch.equal(outs[0][1], outs[2][1]), "dgrad differs between the persistent and the per-tile kernel"
    assert torch.equal(outs[0][0], outs[1][0]), "forward differs between im2col and cp.async"
    assert torch.equal(outs[0][1], outs[1][1]), "dgrad differs between im2col and cp.async"
    assert torch.equal(outs[0][2], outs[1][2]), "wgrad differs between im2col and cp.async"
    dwr = torch.nn.grad.conv2d_weight(_bf(x.cpu()).permute(0, 3, 1, 2), w.shape, _bf(dy.cpu()).permute(0, 3, 1, 2), stride=s, padding=p)
    close(outs[0][2], dwr, 1e-4 if ops._tc_wgrad_ok(geom, "cl", "cl") or True and Cy % 64 == 0 else 1e-2,
          "wgrad vs reference on bf16-rounded operands")
    # and both agree with the fp32 CUDA-core path within the bf16 operand bound
    y32 = ops.conv_forward(geom, ops.WeightPacks(), w, x, "cl", "cl") if not transposed else \
        ops.conv_dgrad(geom, ops.WeightPacks(), w, dy, "cl", (H, H), "cl")
    close(outs[0][0], y32, 2e-2, "tc vs fp32")


# ---------------------------------------------------------------------------------------------------------
# packed GEMM operands follow the weights (round-1 defect: they went stale after raw-pointer / fused optimizer updates)
# ---------------------------------------------------------------------------------------------------------
def _conv_operands(g, w, tc, sources=None):
    packs = ops.WeightPacks(sources)
    fwd = packs.get(("fwd", tc) + g.key(), w, lambda src, rec: ops._pack_fwd(g, src, tc, rec))
    dg = packs.get(("dgrad", tc) + g.key(), w, lambda src, rec: ops._pack_dgrad(g, src, tc, rec))
    return packs, fwd, dg


def _emul_operands(g, w, tc, n_src=1):
    """the same matrices through the CPU restatement of b200_pack_weight"""
    real = _lib.K
    _lib.K = E
    try:
        srcs = tuple(w.cpu().chunk(n_src, 0)) if n_src > 1 else (w.cpu(),)
        fwd = ops._pack_fwd(g, srcs, tc, [])
        dg = ops._pack_dgrad(g, srcs, tc, [])
    finally:
        _lib.K = real
    return fwd, dg


@pytest.mark.parametrize("tc", [False, True])
@pytest.mark.parametrize("geom", [(64, 128, 4, 4, 2, 1, 0, None), (128, 256, 3, 3, 1, 1, 0, None), (64, 256, 5, 5, 1, 2, 128, 192),
                                  (192, 64, 1, 1, 1, 0, 0, None)])
def test_pack_refresh_follows_raw_weight_updates(K, geom, tc):
    """pack once, change the weights WITHOUT moving the tensor version (what b200_adam_multi and torch's fused Adam do),
    refresh_packs() -> every operand (forward + the stride^2 data-gradient phases) equals a fresh pack, bit for bit"""
    Cx, Cy, kh, kw, s, p, off, tot = geom
    g = ops.ConvGeom(Cx, Cy, kh, kw, s, p, off, tot)
    gen = torch.Generator().manual_seed(3)
    w = torch.randn(Cy, g.cx_total, kh, kw, generator=gen).cuda()
    packs, fwd, dg = _conv_operands(g, w, tc)
    v0 = w._version
    w2 = torch.randn(Cy, g.cx_total, kh, kw, generator=gen)
    w.data.copy_(w2)          # `.data` writes do not move the tensor version (nor do raw-pointer kernels)
    assert w._version == v0
    stale = packs.get(("fwd", tc) + g.key(), w, None)          # cache hit: the stamp cannot see the raw update
    assert stale[0] is fwd[0]
    n = ops.refresh_packs([w])
    assert n >= 1 + sum(1 for q in dg if q is not None)
    torch.cuda.synchronize()
    ref_fwd, ref_dg = _emul_operands(g, w2, tc)
    assert torch.equal(fwd[0].cpu(), ref_fwd[0]) and fwd[1] == ref_fwd[1]
    for a, r in zip(dg, ref_dg):
        assert (a is None) == (r is None)
        if a is not None:
            assert torch.equal(a[0].cpu(), r[0])
    # a version-visible update (load_state_dict, foreach optimizers) re-packs lazily into the SAME buffers
    with torch.no_grad():
        w.mul_(2.0)
    again = packs.get(("fwd", tc) + g.key(), w, lambda src, rec: ops._pack_fwd(g, src, tc, rec))
    assert again[0].data_ptr() == fwd[0].data_ptr()
    assert torch.equal(again[0].cpu(), _emul_operands(g, w2 * 2.0, tc)[0][0])


@pytest.mark.parametrize("tc", [False, True])
def test_pack_from_two_parameters(K, tc):
    """SPADE's fused gamma|beta convolution (normalization.py:101-104): the 2C-output operand is packed from the two
    parameters directly (row slices forward, channel slices in the data-gradient matrices)"""
    g = ops.ConvGeom(128, 256, 3, 3, 1, 1)
    gen = torch.Generator().manual_seed(4)
    wg, wb = torch.randn(128, 128, 3, 3, generator=gen).cuda(), torch.randn(128, 128, 3, 3, generator=gen).cuda()
    w = torch.cat([wg, wb])
    packs, fwd, dg = _conv_operands(g, w, tc, sources=(wg, wb))
    one, fwd1, dg1 = _conv_operands(g, w, tc)
    assert torch.equal(fwd[0], fwd1[0])
    for a, r in zip(dg, dg1):
        assert torch.equal(a[0], r[0])
    wb.data.mul_(-3.0)
    ops.refresh_packs([wb])
    ref_fwd, ref_dg = _emul_operands(g, torch.cat([wg, wb]), tc)
    assert torch.equal(fwd[0].cpu(), ref_fwd[0])
    for a, r in zip(dg, ref_dg):
        assert torch.equal(a[0].cpu(), r[0])


# ---------------------------------------------------------------------------------------------------------
# tcgen05 kind::tf32 — fp32 tensors on the tensor cores (BASELINE config 2, "fp32" half)
# ---------------------------------------------------------------------------------------------------------
TF32_CASES = [c for c in CONV_CASES if c[7] == "cl" and c[8] == "cl" and c[0] % 64 == 0 and c[1] % 64 == 0] + [
    (64, 64, 3, 1, 1, 64, 4, "cl", "cl", True), (512, 640, 1, 1, 0, 4, 3, "cl", "cl", False)]


@pytest.mark.parametrize("case", TF32_CASES)
def test_conv_tf32_fwd_bwd(K, case):
    """forward and data gradient on the kind::tf32 kernels (TMA TFLOAT32 operands, fp32 accumulation in TMEM), weight
    gradient as three bf16 tensor-core GEMMs over two-term bf16 splits of the fp32 operands, against F.conv2d autograd in fp32.  Stated bound: 2e-3 relative per op (10-bit operand mantissas; SURVEY.md App.
    D); and against the same computation on tf32-ROUNDED operands the kernels must agree to 3e-4 (rounding mode of the TMA
    conversion + summation order), which proves the error is operand rounding, not arithmetic."""
    Cx, Cy, k, s, p, H, N, xl, ol, has_b = case
    g = torch.Generator().manual_seed(Cx * 5 + Cy + k)
    x = torch.randn(N, Cx, H, H, generator=g)
    w = torch.randn(Cy, Cx, k, k, generator=g) / (Cx * k * k) ** 0.5
    b = torch.randn(Cy, generator=g) if has_b else None
    ops.set_precision("tf32")
    try:
        geom = ops.ConvGeom(Cx, Cy, k, k, s, p)
        assert ops._tc_fwd_ok(geom, xl) == ops.TC_TF32 and ops._tc_wgrad_ok(geom, xl, ol) == ops.TC_TF32
        xd = _to_layout(x, xl).cuda().requires_grad_(True)
        wd = w.cuda().requires_grad_(True)
        bd = b.cuda().requires_grad_(True) if has_b else None
        before = K.launch_count()
        y = ops.conv2d(xd, wd, bd, geom, ops.WeightPacks(), xl, ol, relu=False)
        assert y.dtype == torch.float32
        xr, wr = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
        br = b.clone().requires_grad_(True) if has_b else None
        yr = _conv_ref(xr, wr, br, s, p, False, None)
        gy = torch.randn(yr.shape, generator=g)
        yr.backward(gy)
        close(_from_layout(y, ol), yr, 2e-3, "tf32 fwd")
        y.backward(_to_layout(gy, ol).cuda())
        close(_from_layout(xd.grad, xl), xr.grad, 2e-3, "tf32 dgrad")
        close(wd.grad, wr.grad, 1e-4, "tf32-mode wgrad (three bf16 GEMMs over two-term splits: tighter than tf32)")
        if has_b:
            close(bd.grad, br.grad, 1e-5, "bias grad")
        from abi_emul import _tf32
        yq = _conv_ref(_tf32(x), _tf32(w), b, s, p, False, None)
        close(_from_layout(y, ol), yq, 3e-4, "tf32 fwd vs rounded operands")
        dxq = torch.nn.grad.conv2d_input(x.shape, _tf32(w), _tf32(gy), stride=s, padding=p)
        close(_from_layout(xd.grad, xl), dxq, 3e-4, "tf32 dgrad vs rounded operands")
    finally:
        ops.set_precision("fp32")


def test_tf32_transpose_sn_and_lstm(K):
    """the other users of the GEMM kernels in tf32 mode: ConvTranspose2d (dgrad phases as forward), batched spectral-norm
    convolution (grouped weight gradient through W / sigma_g) and the ConvLSTM (hoisted + recurrent GEMMs, BPTT)"""
    g = torch.Generator().manual_seed(21)
    ops.set_precision("tf32")
    try:
        x = torch.randn(2, 256, 8, 8, generator=g)
        w = torch.randn(256, 128, 4, 4, generator=g) / (256 * 4) ** 0.5
        geom = ops.ConvGeom(128, 256, 4, 4, 2, 1)
        xd, wd = _to_layout(x, "cl").cuda().requires_grad_(True), w.cuda().requires_grad_(True)
        y = ops.conv_transpose2d(xd, wd, geom, ops.WeightPacks(), (16, 16))
        xr, wr = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
        yr = F.conv_transpose2d(xr, wr, None, stride=2, padding=1)
        gy = torch.randn(yr.shape, generator=g)
        yr.backward(gy)
        y.backward(_to_layout(gy, "cl").cuda())
        close(_from_layout(y, "cl"), yr, 2e-3, "convT fwd")
        close(_from_layout(xd.grad, "cl"), xr.grad, 2e-3, "convT dgrad")
        close(wd.grad, wr.grad, 2e-3, "convT wgrad")
    finally:
        ops.set_precision("fp32")
    # ConvLSTM: 2 layers, ragged sequences
    o2i = torch.tensor([0, 0, 0, 1, 2, 2])
    x = torch.randn(6, 64, 8, 8, generator=g) * 0.5
    sd, layers, params, cin = {}, [], [], 64
    for i, hid in enumerate((64, 32)):
        sd["c.cell_list.%d.conv.weight" % i] = (torch.randn(4 * hid, cin + hid, 5, 5, generator=g) / ((cin + hid) * 25) ** 0.5).requires_grad_(True)
        sd["c.cell_list.%d.conv.bias" % i] = (torch.randn(4 * hid, generator=g) * 0.1).requires_grad_(True)
        layers.append(ops.ConvLSTMLayer(cin, hid, 5))
        cin = hid
    xr = x.clone().requires_grad_(True)
    ref = O.conv_lstm(sd, "c", xr, o2i, hidden=(64, 32))
    gy = torch.randn(ref.shape, generator=g)
    ref.backward(gy)
    ops.set_precision("tf32")
    try:
        pd = []
        for i in range(2):
            pd += [sd["c.cell_list.%d.conv.weight" % i].detach().cuda().requires_grad_(True),
                   sd["c.cell_list.%d.conv.bias" % i].detach().cuda().requires_grad_(True)]
        xd = x.permute(0, 2, 3, 1).contiguous().cuda().requires_grad_(True)
        out = ops.conv_lstm(xd, ops.get_plan(o2i, 3, "cuda"), layers, pd)
        out.backward(gy.permute(0, 2, 3, 1).contiguous().cuda())
        close(out.permute(0, 3, 1, 2), ref, 2e-3, "tf32 convlstm fwd")
        close(xd.grad.permute(0, 3, 1, 2), xr.grad, 4e-3, "tf32 convlstm dx")
        for i in range(2):
            close(pd[2 * i].grad, sd["c.cell_list.%d.conv.weight" % i].grad, 4e-3, "tf32 convlstm dw%d" % i)
    finally:
        ops.set_precision("fp32")


# ---------------------------------------------------------------------------------------------------------
# fused step arithmetic (loss.cu; train64.py:195-252, 284-364) against torch's functional losses + autograd
# ---------------------------------------------------------------------------------------------------------
def test_fused_loss_terms_match_torch(K):
    g = torch.Generator().manual_seed(17)
    O_, N, A, C = 37, 5, 106, 179
    dev = "cuda"
    lam = dict(adv=1.5, cls=0.7, att=2.0, rec=1.3, z=8.0, kl=0.01)
    src4 = torch.randn(4 * N, generator=g) * 2
    cls3 = torch.randn(3 * O_, C, generator=g) * 3
    att3 = torch.randn(3 * O_, A, generator=g) * 2
    labels = torch.randint(0, C, (O_,), generator=g)
    tgt = (torch.rand(O_, A, generator=g) < 0.02).float()
    tgt[::3] = 0                                                   # un-annotated objects
    sel = (tgt.sum(1) != 0).float()
    pw = torch.rand(A, generator=g) * 50 + 1
    img_a, img_b = torch.randn(N, 3, 16, 16, generator=g), torch.randn(N, 3, 16, 16, generator=g)
    mask = torch.tensor([0.0, 1.0, 1.0, 1.0, 1.0])
    mu2, z = torch.randn(2 * O_, 64, generator=g), torch.randn(O_, 64, generator=g)
    mu, lv = torch.randn(O_, 64, generator=g), torch.randn(O_, 64, generator=g) * 0.3
    w3, w4 = (0.4, 0.4, 0.2), (0.4, 0.4, 0.2, 1.0)

    def leaf(t, d):
        return t.clone().to(d).requires_grad_(True)

    # ---- reference: the PyTorch formulation of b200gan/step.py (d_loss_torch / g_loss_torch) ----
    r = {k: leaf(v, "cpu") for k, v in dict(src4=src4, cls3=cls3, att3=att3, img=img_a, mu2=mu2, mu=mu, lv=lv).items()}
    bce = F.binary_cross_entropy_with_logits
    per = bce(r["src4"], torch.tensor([0.0, 0, 0, 1]).repeat_interleave(N), reduction="none").view(4, N).mean(1)
    t_fake, t_real = lam["adv"] * (torch.tensor(w3) * per[:3]).sum(), lam["adv"] * per[3]
    t_cls = lam["cls"] * (torch.tensor(w3) * F.cross_entropy(r["cls3"], labels.repeat(3), reduction="none").view(3, O_).mean(1)).sum()
    idx = sel.nonzero().view(-1)
    t_att = lam["att"] * sum(w * bce(r["att3"][i * O_:(i + 1) * O_].index_select(0, idx), tgt.index_select(0, idx), pos_weight=pw)
                             for i, w in enumerate(w3))
    t_rec = lam["rec"] * (mask * (r["img"] - img_b).abs().view(N, -1).mean(1)).sum() / 4.0
    t_z = lam["z"] * (0.5 * (r["mu2"][:O_] - z).abs().mean() + 0.5 * (r["mu2"][O_:] - z).abs().mean())
    t_kl = lam["kl"] * -0.5 * torch.sum(1 + r["lv"] - r["mu"].pow(2) - r["lv"].exp())
    want = [t_fake, t_real, t_cls, t_att, t_rec, t_z, t_kl]
    sum(want).backward()

    # ---- kernels ----
    d = {k: leaf(v, dev) for k, v in dict(src4=src4, cls3=cls3, att3=att3, img=img_a, mu2=mu2, mu=mu, lv=lv).items()}
    acc = ops.FusedLoss(["fake", "real", "cls", "att", "rec", "z", "kl"], dev)
    acc.add_bce_groups("fake", d["src4"], 4, (0, 0, 0, 1), w4, lam["adv"], split_group=3)
    acc.add_ce_groups("cls", d["cls3"], labels.to(dev), 3, w3, lam["cls"])
    acc.add_bce_pos_weight_rows("att", d["att3"], tgt.to(dev), sel.to(dev), int(sel.sum()), pw.to(dev), 3, w3, lam["att"])
    acc.add_l1_rows("rec", d["img"], img_b.to(dev), N, mask.to(dev), 4.0, lam["rec"])
    acc.add_l1_rows("z", d["mu2"], z.to(dev), 2, None, 1.0, 0.5 * lam["z"], broadcast_b=True)
    acc.add_kl("kl", d["mu"], d["lv"], lam["kl"])
    total = acc.total()
    total.backward()
    terms = acc.terms.cpu()
    for i, t in enumerate(want):
        assert abs(float(terms[i]) - float(t)) <= 2e-6 * max(1.0, abs(float(t))), (acc.names[i], float(terms[i]), float(t))
    assert abs(float(total) - float(sum(want))) <= 2e-6 * abs(float(sum(want)))
    for k in r:
        close(d[k].grad, r[k].grad, 2e-6, "fused loss grad " + k)
    # rows without annotation and groups with zero weight receive exactly zero gradient
    assert float(d["att3"].grad.view(3, O_, A)[:, sel == 0].abs().max()) == 0.0
    acc2 = ops.FusedLoss(["c"], dev)
    c4 = leaf(torch.randn(4 * O_, C, generator=g), dev)
    acc2.add_ce_groups("c", c4, labels.to(dev), 4, (0, 0, 0, 1), 1.0)
    acc2.total().backward()
    assert float(c4.grad[:3 * O_].abs().max()) == 0.0 and float(c4.grad[3 * O_:].abs().max()) > 0.0
    want_c = F.cross_entropy(c4.detach().cpu()[3 * O_:], labels)
    assert abs(float(acc2.terms[0]) - float(want_c)) < 2e-6 * float(want_c)


@pytest.mark.parametrize("N,H,S,per", [(4, 64, 32, 8), (3, 128, 64, 30), (2, 48, 16, 5)])
def test_crop_staged_kernels_bit_identical_to_elementwise(K, N, H, S, per):
    """the shared-memory staged crop kernels (footprint / gradient tiles staged per block, coordinates computed once per box)
    keep the arithmetic and the summation order of the element-wise kernels: forward and backward agree bit for bit, incl.
    boxes touching / leaving the image, a full-image box (footprint larger than the staging tile at 128x128) and a
    degenerate box"""
    g = torch.Generator().manual_seed(N * H + S)
    B = N * per
    feats = torch.randn(N, 3, H, H, generator=g).cuda()
    xy0 = torch.rand(B, 2, generator=g) * 0.7
    boxes = torch.cat([xy0, (xy0 + torch.rand(B, 2, generator=g) * 0.5 + 0.02)], 1)
    boxes[0] = torch.tensor([0.0, 0.0, 1.0, 1.0])
    boxes[1] = torch.tensor([0.5, 0.5, 0.5, 0.5])
    boxes[2] = torch.tensor([-0.2, 0.1, 0.3, 1.4])
    boxes[3] = torch.tensor([0.9, 0.8, 0.1, 0.2])              # inverted
    boxes = boxes.cuda()
    b2i = torch.sort(torch.randint(0, N, (B,), generator=g))[0]
    plan = ops.get_plan(b2i, N, "cuda")
    w = ops.crop_weights(S, "cuda")
    gy = torch.randn(B, 3, S, S, generator=g).cuda()
    outs = {}
    for staged in (True, False):
        prev = K.crop_set_staged(staged)
        try:
            outs[staged] = (K.crop_fwd(feats, boxes, plan.box_to_img, w, w, S, S),
                            K.crop_bwd(gy, boxes, plan.img_box_start, plan.box_order, w, w, N, H, H))
        finally:
            K.crop_set_staged(prev)
    assert torch.equal(outs[True][0], outs[False][0])
    assert torch.equal(outs[True][1], outs[False][1])
    assert torch.isfinite(outs[True][0]).all() and float(outs[True][1].abs().max()) > 0


# ---------------------------------------------------------------------------------------------------------
# batch-norm statistics accumulated in the convolution epilogue (b200_conv_desc.col_stats + b200_bn_stats_slabs)
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("precision", ["bf16", "tf32"])
@pytest.mark.parametrize("case", [(64, 64, 3, 1, 1, 32, 6, 3), (64, 128, 4, 2, 1, 32, 6, 3), (128, 192, 3, 1, 1, 8, 7, 1),
                                  (256, 512, 4, 2, 1, 16, 4, 2), (3, 64, 7, 1, 3, 32, 6, 2)])
def test_conv_epilogue_statistics_match_the_statistics_pass(K, case, precision):
    """a convolution asked for `stats` hands the normalisation the per-slab (sum, sum of squares) pairs of its STORED output:
    mean / biased variance / running statistics from b200_bn_stats_slabs equal those of the separate statistics pass over the
    stored tensor (b200_bn_stats) to summation order, for grouped calls, partial last tiles (M % 128 != 0) and padded channel
    tiles; the convolution output itself is unchanged bit for bit"""
    Cx, Cy, k, s, p, H, N, groups = case
    if precision == "tf32" and Cx < 32:
        pytest.skip("3-channel layers run on the CUDA-core kernels in tf32 mode")
    g = torch.Generator().manual_seed(Cx + Cy + H)
    xl = "nchw" if Cx == 3 else "cl"
    x = torch.randn(N, Cx, H, H, generator=g)
    w = torch.randn(Cy, Cx, k, k, generator=g) / (Cx * k * k) ** 0.5
    ops.set_precision(precision)
    prev = ops.FUSE_BN_STATS
    ops.FUSE_BN_STATS = True            # (off by default: measured slower inside the step — an experiment kept correct)
    try:
        geom = ops.ConvGeom(Cx, Cy, k, k, s, p)
        xd = _to_layout(x, xl).cuda()
        wd = w.cuda()
        y = ops.conv2d(xd, wd, None, geom, ops.WeightPacks(), xl, "cl", stats=True)
        cs = getattr(y, "_b200_colstats", None)
        rows = y.shape[0] * y.shape[1] * y.shape[2]
        y0 = ops.conv2d(xd, wd, None, geom, ops.WeightPacks(), xl, "cl", stats=False)
        assert torch.equal(y, y0)
        if cs is None:
            pytest.skip("this launch does not run on the persistent kernel (split-K): the statistics pass is used")
        assert cs.shape[0] == (rows + 127) // 128 * 4
        if (rows // groups) % 32:
            pytest.skip("rows per group not slab aligned: the statistics pass is used")
        rm_a, rv_a = torch.zeros(Cy, device="cuda"), torch.ones(Cy, device="cuda")
        rm_b, rv_b = rm_a.clone(), rv_a.clone()
        m_a, v_a = K.bn_stats_slabs(cs, rows, Cy, rm_a, rv_a, 0.1, groups)
        m_b, v_b = K.bn_stats(y.reshape(rows, Cy), rm_b, rv_b, 0.1, groups)
        assert float((m_a - m_b).abs().max()) <= 1e-6 * max(1.0, float(m_b.abs().max()))
        close(v_a, v_b, 1e-5, "variance")
        close(rm_a, rm_b, 1e-5, "running mean")
        close(rv_a, rv_b, 1e-5, "running var")
    finally:
        ops.FUSE_BN_STATS = prev
        ops.set_precision("fp32")


# ---- pooled convolution: avg_pool2(conv(x)) as one stride-2 convolution with the folded weight ------------------------------
@pytest.mark.parametrize("kh,kw", [(3, 3), (1, 1), (5, 3)])
def test_fold_pool_weight_matches_definition(K, kh, kw):
    """b200_fold_pool_weight against the definition in include/b200gan.h (the emulation), and the defining identity itself:
    conv_{(kh+1)x(kw+1), stride 2}(x; W4) == avg_pool2(conv_{kh x kw}(x; W)) in torch fp64"""
    g = torch.Generator().manual_seed(kh * 10 + kw)
    w = torch.randn(6, 5, kh, kw, generator=g)
    w4 = torch.empty(6, 5, kh + 1, kw + 1, device="cuda")
    K.fold_pool_weight(w.cuda(), w4)
    want = E.fold_pool_weight(w, torch.empty(6, 5, kh + 1, kw + 1))
    close(w4.cpu(), want, 1e-7, "folded weight")
    if kh == kw:
        p = kh // 2
        x = torch.randn(2, 5, 9, 12, generator=g).double()
        a = F.avg_pool2d(F.conv2d(x, w.double(), padding=p), 2)
        b = F.conv2d(x, want.double(), stride=2, padding=p)
        close(b, a, 1e-6, "fold identity")          # (W4 above is the fp32 image)


@pytest.mark.parametrize("precision", ["fp32", "bf16", "tf32"])
@pytest.mark.parametrize("case", [(64, 128, 3, 1, 16, 3, 16), (64, 64, 3, 1, 32, 3, 2), (128, 256, 3, 1, 8, 1, 4),
                                  (64, 128, 1, 0, 16, 2, 8), (128, 128, 3, 1, 6, 3, 2)])
def test_pooled_conv_equals_conv_then_pool(K, case, precision):
    """Conv2d(pool=True) semantics: forward, masked data gradient, spectral-norm weight gradient (grouped finish and the
    per-call fallback) and bias gradient against torch's avg_pool2d(conv2d(x, W / sigma_g) + b) per batched call"""
    Cx, Cy, k, p, H, groups, n = case
    g = torch.Generator().manual_seed(Cx + Cy + H + groups)
    x = torch.relu(torch.randn(groups * n, Cx, H, H, generator=g))          # a ReLU output: exercises mask_input_grad
    w = torch.randn(Cy, Cx, k, k, generator=g) / (Cx * k * k) ** 0.5
    b = torch.randn(Cy, generator=g)
    u = F.normalize(torch.randn(Cy, generator=g), dim=0)
    v = F.normalize(torch.randn(Cx * k * k, generator=g), dim=0)
    gy = torch.randn(groups * n, Cy, H // 2, H // 2, generator=g)
    tol = {"fp32": 3e-5, "tf32": 3e-3, "bf16": 2e-2}[precision]
    ops.set_precision(precision)
    try:
        act = ops.act_dtype()
        xd = _to_layout(x, "cl").cuda().to(act).requires_grad_(True)
        wd, bd, ud, vd = w.cuda().requires_grad_(True), b.cuda().requires_grad_(True), u.cuda(), v.cuda()
        geom = ops.ConvGeom.pooled(Cx, Cy, k, k, p)
        packs = ops.WeightPacks()
        sn = ops.sn_iterate(wd.detach(), ud, vd, groups, True)
        y = ops.conv2d(xd, wd, bd, geom, packs, "cl", "cl", sn=sn, mask_input_grad=True)
        assert tuple(y.shape) == (groups * n, H // 2, H // 2, Cy)
        y.backward(_to_layout(gy, "cl").cuda().to(y.dtype))
        # the folded operands follow a raw weight update through refresh_packs (the optimizer hook's path)
        wd.data.mul_(0.5)
        ops.refresh_packs([wd])
        y2 = ops.conv2d(xd.detach(), wd.detach(), None, geom, packs, "cl", "cl")
    finally:
        ops.set_precision("fp32")
    st = {"l.weight_orig": w.clone().requires_grad_(True), "l.weight_u": u.clone(), "l.weight_v": v.clone()}
    bias = b.clone().requires_grad_(True)
    xr = x.clone().requires_grad_(True)
    ys = []
    for gi in range(groups):
        wn = O.sn_weight(st, "l", True)
        ys.append(F.avg_pool2d(F.conv2d(xr[gi * n:(gi + 1) * n], wn, bias, padding=p), 2))
    yr = torch.cat(ys)
    yr.backward(gy)
    close(_from_layout(y.float(), "cl"), yr, tol, "pooled conv fwd")
    close(_from_layout(xd.grad.float(), "cl"), xr.grad * (x > 0), tol, "pooled conv masked dgrad")
    close(wd.grad, st["l.weight_orig"].grad, tol if precision != "fp32" else 6e-5, "pooled conv sn wgrad")
    close(bd.grad, bias.grad, 2e-5 if precision == "fp32" else tol, "pooled conv bias grad")
    want2 = F.avg_pool2d(F.conv2d(x, 0.5 * w, None, padding=p), 2)
    close(_from_layout(y2.float(), "cl"), want2, tol, "pooled conv after refresh_packs")
