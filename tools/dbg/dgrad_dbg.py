import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "multiple-objects-gan_b200")]
import torch, torch.nn.functional as F
from mog_b200 import ops
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
def run(N, H, Ci, Co, k, s, p, prec, scale=1.0):
    ops.set_precision(prec)
    torch.manual_seed(0)
    x = torch.randn(N, Ci, H, H, device="cuda")
    w = torch.randn(Co, Ci, k, k, device="cuda") / (Ci * k * k) ** 0.5
    xr = x.double().requires_grad_(True); wr = w.double().requires_grad_(True)
    yr = F.conv2d(xr, wr, None, s, p)
    g = torch.randn_like(yr) * scale
    g = g - g.mean(dim=(0, 2, 3), keepdim=True) + 0.0
    yr.backward(g)
    xm = x.permute(0, 2, 3, 1).contiguous().requires_grad_(True)
    wm = w.clone().requires_grad_(True)
    ym = ops.conv2d(xm, wm, None, s, p)
    ym.backward(g.float().permute(0, 2, 3, 1).contiguous())
    dx = xm.grad.permute(0, 3, 1, 2).double()
    e = dx - xr.grad
    rel = float(e.norm() / xr.grad.norm())
    relw = float((wm.grad.double() - wr.grad).norm() / wr.grad.norm())
    rely = float((ym.permute(0, 3, 1, 2).double() - yr).norm() / yr.norm())
    print("%s N%d H%d %d->%d k%d s%d p%d: y %.2e dx %.2e dw %.2e | dx err mean/ch %s" % (prec, N, H, Ci, Co, k, s, p, rely, rel, relw,
          ["%.1e" % v for v in e.mean(dim=(0, 2, 3))[:4].tolist()]))
for prec in ("bf16x3",):
    run(64, 4, 256, 128, 3, 1, 1, prec)
    run(31, 4, 1024, 768, 3, 1, 1, prec)
    run(64, 8, 128, 256, 4, 2, 1, prec)
    run(6, 4, 64, 64, 3, 1, 1, prec)
    run(64, 4, 3072, 1536, 3, 1, 1, prec)
    run(64, 8, 1536, 3072, 4, 2, 1, prec)
