import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "multiple-objects-gan_b200")]
import torch, torch.nn as nn, torch.nn.functional as F
from mog_b200 import ops
from mog_b200.ops import ACT_LRELU
def run(prec, S, Mimg, HW, C, offset=0.0, scale=1.0):
    ops.set_precision(prec)
    torch.manual_seed(1)
    N = S * Mimg
    x = (torch.randn(N, HW, HW, C, device="cuda") * scale + offset)
    g = torch.randn(N, HW, HW, C, device="cuda")
    bn = nn.BatchNorm2d(C).cuda().train()
    with torch.no_grad():
        bn.weight.normal_(1, 0.1); bn.bias.normal_(0, 0.1)
    xm = x.clone().requires_grad_(True)
    y = ops.bn_act(xm, bn, ACT_LRELU, segments=S)
    y.backward(g)
    dgm, dbm = bn.weight.grad.clone(), bn.bias.grad.clone()
    bn.weight.grad = None; bn.bias.grad = None
    # torch double, per segment
    xd = x.double().permute(0, 3, 1, 2).requires_grad_(True)
    w = bn.weight.detach().double().requires_grad_(True); b = bn.bias.detach().double().requires_grad_(True)
    outs = []
    for s in range(S):
        outs.append(F.leaky_relu(F.batch_norm(xd[s * Mimg:(s + 1) * Mimg], None, None, w, b, True, 0.1, 1e-5), 0.2))
    yd = torch.cat(outs, 0)
    yd.backward(g.double().permute(0, 3, 1, 2))
    rel = lambda a, r: float((a.double() - r).norm() / r.norm())
    dxr = xd.grad.permute(0, 2, 3, 1)
    msg = "%s S%d M%d C%d off %.0f sc %.0e: y %.1e dx %.1e (seg0 %.1e seg1 %.1e) dgamma %.1e dbeta %.1e" % (
        prec, S, Mimg * HW * HW, C, offset, scale, rel(y, yd.permute(0, 2, 3, 1)), rel(xm.grad, dxr), rel(xm.grad[:Mimg], dxr[:Mimg]),
        rel(xm.grad[Mimg:], dxr[Mimg:]) if S > 1 else 0, rel(dgm, w.grad), rel(dbm, b.grad))
    pl = getattr(xm.grad, "_mog_planes", None)
    print(msg, "planes" if pl is not None else "")
for prec in ("fp32", "bf16x3"):
    run(prec, 2, 4, 4, 128)
    run(prec, 2, 4, 4, 128, 30.0)
    run(prec, 1, 8, 4, 128, 30.0)
    run(prec, 2, 4, 32, 32, 30.0)
    run(prec, 2, 4, 32, 32, 300.0)
    run(prec, 2, 4, 4, 128, 0.0, 1e-3)
