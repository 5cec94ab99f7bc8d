"""probe: where does the tf32 weight-gradient kernel put a single product?  R[m, c] = sum_pix P[pix, m] * G[pix, c]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "attribute-guided-image-generation-from-layout_b200")):
    sys.path.insert(0, p)
import torch
from b200gan import _lib, ops
from b200gan._lib import ConvDesc
K = _lib.Kernels()
def run(Cin, Cout, Q, P, G, kind):
    ws = torch.zeros(Cout * Cin, device="cuda")
    xs = (Q * Cin, Q * Cin, Cin, 1); ds = (Q * Cout, Q * Cout, Cout, 1)
    d = ConvDesc(B=1, Qh=1, Qw=Q, Cin=Cin, Cout=Cout, Th=1, Tw=1, in_sy=1, in_sx=1, tap_sy=1, tap_sx=1, tap_oy=0, tap_ox=0, Hi=1, Wi=Q,
                 up_shift=0, in_sn=xs[0], in_sh=xs[1], in_sw=xs[2], in_sc=1, out_sy=1, out_sx=1, out_oy=0, out_ox=0, Ho=1, Wo=Q,
                 out_sn=ds[0], out_sh=ds[1], out_sw=ds[2], out_sc=1, ldw=Cin, relu=0)
    K.wgrad_gemm(d, P, G, ws, 1, kind)
    torch.cuda.synchronize()
    return ws.view(Cout, Cin).cpu()
for (Cin, Cout, Q) in ((32, 32, 64), (64, 64, 64), (128, 128, 128)):
    for (p0, m0, c0) in ((0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1), (9, 5, 3), (17, 33 % Cout, 40 % Cin), (63, Cout - 1, Cin - 1)):
        P = torch.zeros(Q, Cout); G = torch.zeros(Q, Cin)
        P[p0, m0] = 2.0; G[p0, c0] = 3.0
        for kind, cast in ((2, lambda t: t.cuda()), (1, lambda t: t.cuda().bfloat16())):
            if kind == 1 and (Cin % 64 or Cout % 64):
                continue
            R = run(Cin, Cout, Q, cast(P), cast(G), kind)
            nz = R.nonzero().tolist()
            print("C=%d/%d Q=%d kind=%d  product at pix %d (m %d, c %d) -> nonzeros %s vals %s" %
                  (Cin, Cout, Q, kind, p0, m0, c0, nz[:6], [float(R[a, b]) for a, b in nz[:6]]))
    # all-ones operands: every entry must be Q
    R = run(Cin, Cout, Q, torch.ones(Q, Cout).cuda(), torch.ones(Q, Cin).cuda(), 2)
    print("ones C=%d/%d Q=%d tf32: min %.1f max %.1f (expected %d) nonzero frac %.3f" % (Cin, Cout, Q, float(R.min()), float(R.max()), Q, float((R != 0).float().mean())))
    g = torch.Generator().manual_seed(0)
    P = torch.randn(Q, Cout, generator=g); G = torch.randn(Q, Cin, generator=g)
    want = P.t() @ G
    for kind, cast in ((2, lambda t: t.cuda()), (1, lambda t: t.cuda().bfloat16())):
        if kind == 1 and (Cin % 64 or Cout % 64):
            continue
        R = run(Cin, Cout, Q, cast(P), cast(G), kind)
        print("random C=%d/%d Q=%d kind=%d rel err %.3e   (vs transposed: %.3e)" %
              (Cin, Cout, Q, kind, float((R - want).norm() / want.norm()), float((R - want.t()).norm() / want.norm()) if Cin == Cout else -1))
