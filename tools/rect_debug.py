import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "multiple-objects-gan_b200")]
import torch, torch.nn.functional as F
from mog_b200 import ops
torch.backends.cudnn.allow_tf32 = False
def run(ks, pad, stride, H, W, Ci, Co, prec="bf16x3", relu=True):
    torch.manual_seed(5)
    x = torch.randn(2, H, W, Ci, device="cuda")
    w = torch.randn(Co, Ci, ks[0], ks[1], device="cuda") / (Ci * ks[0] * ks[1]) ** 0.5
    xr = x.permute(0, 3, 1, 2).detach().clone().requires_grad_(True)
    xm = x.clone().requires_grad_(True)
    y = ops.conv2d(xm, w, None, stride, pad if pad[0] != pad[1] else pad[0], False, ops.ACT_RELU if relu else 0, ops.PREC_NAMES[prec])
    yr = F.conv2d(xr, w, None, stride, pad)
    if relu: yr = F.relu(yr)
    g = torch.randn_like(y)
    y.backward(g); yr.backward(g.permute(0, 3, 1, 2))
    e = (xm.grad - xr.grad.permute(0, 2, 3, 1))
    ref = xr.grad.permute(0, 2, 3, 1)
    print(ks, pad, stride, H, Ci, Co, "relu" if relu else "lin", "fwd rel %.2e" % ((y - yr.permute(0,2,3,1)).norm() / yr.norm()).item(),
          "dgrad rel %.2e" % (e.norm() / ref.norm()).item())
    er = e.pow(2).sum((0, 2, 3)).sqrt() / ref.pow(2).sum((0,2,3)).sqrt().clamp_min(1e-20)
    ec = e.pow(2).sum((0, 1, 2)).sqrt() / ref.pow(2).sum((0,1,2)).sqrt().clamp_min(1e-20)
    print("  rows worst:", [(int(i), "%.1e" % er[i].item()) for i in er.argsort(descending=True)[:6]], " median %.1e" % er.median().item())
    print("  chans worst:", [(int(i), "%.1e" % ec[i].item()) for i in ec.argsort(descending=True)[:6]], " median %.1e" % ec.median().item())
for relu in (True, False):
    run((3, 3), (0, 0), 1, 73, 73, 80, 192, relu=relu)
    run((1, 7), (0, 3), 1, 17, 17, 128, 192, relu=relu)
    run((3, 3), (0, 0), 2, 299, 299, 3, 32, relu=relu)
run((3, 3), (1, 1), 1, 73, 73, 80, 192)
run((3, 3), (0, 0), 1, 72, 72, 80, 192)
# pool
x = torch.randn(3, 17, 17, 24, device="cuda"); xr = x.permute(0,3,1,2).detach().clone().requires_grad_(True); xm = x.clone().requires_grad_(True)
y = ops.avg_pool2d(xm, 3, 1, 1); yr = F.avg_pool2d(xr, 3, 1, 1); g = torch.randn_like(y); y.backward(g); yr.backward(g.permute(0,3,1,2))
d = (xm.grad.permute(0,3,1,2) - xr.grad).abs(); print("avgpool bwd maxdiff", d.max().item(), "fwd", (y.permute(0,3,1,2)-yr).abs().max().item(), d.flatten().argmax().item())
