import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "multiple-objects-gan_b200")]
import torch, torch.nn.functional as F
from mog_b200 import ops
from mog_b200._lib import call
st = lambda: torch.cuda.current_stream().cuda_stream
torch.manual_seed(0)
N,H,W,C = 3,17,17,24
g = torch.randn(N,H,W,C, device="cuda")
dx1 = torch.empty_like(g); dx2 = torch.empty_like(g)
call("mog_pool2d_bwd", None, g.data_ptr(), dx1.data_ptr(), N,H,W,C,3,1,1,1, st())
call("mog_pool2d_bwd", None, g.data_ptr(), dx2.data_ptr(), N,H,W,C,3,1,1,1, st())
ref = F.avg_pool2d(g.permute(0,3,1,2), 3, 1, 1).permute(0,2,3,1)   # symmetric: bwd of avg(3,1,1) == avg(3,1,1) of dy
print("direct call: dx1 vs dx2", (dx1-dx2).abs().max().item(), " vs ref", (dx1-ref).abs().max().item())
# through autograd
xm = torch.randn(N,H,W,C, device="cuda", requires_grad=True)
y = ops.avg_pool2d(xm, 3, 1, 1)
y.backward(g)
print("autograd: vs ref", (xm.grad-ref).abs().max().item())
xr = xm.detach().permute(0,3,1,2).clone().requires_grad_(True)
yr = F.avg_pool2d(xr, 3, 1, 1); yr.backward(g.permute(0,3,1,2))
print("torch bwd vs ref", (xr.grad.permute(0,2,3,1)-ref).abs().max().item())
xr2 = xm.detach().permute(0,3,1,2).contiguous().requires_grad_(True)
yr2 = F.avg_pool2d(xr2, 3, 1, 1); yr2.backward(g.permute(0,3,1,2).contiguous())
print("torch bwd (NCHW contiguous) vs ref", (xr2.grad.permute(0,2,3,1)-ref).abs().max().item())
# relu conv
torch.backends.cudnn.allow_tf32 = False
x = torch.randn(2,73,73,80, device="cuda"); w = torch.randn(192,80,3,3, device="cuda")/27
xr = x.permute(0,3,1,2).detach().clone().requires_grad_(True)
yr = F.relu(F.conv2d(xr, w)); gg = torch.randn(2,71,71,192, device="cuda"); yr.backward(gg.permute(0,3,1,2))
xc = x.permute(0,3,1,2).contiguous().requires_grad_(True)
yc = F.relu(F.conv2d(xc, w)); yc.backward(gg.permute(0,3,1,2).contiguous())
print("torch conv+relu bwd channels_last vs contiguous:", ((xr.grad - xc.grad).norm()/xc.grad.norm()).item())
xm = x.clone().requires_grad_(True)
ops.set_precision("bf16x3")
y = ops.conv2d(xm, w, None, 1, 0, False, ops.ACT_RELU); y.backward(gg)
print("mog vs torch-contiguous:", ((xm.grad.permute(0,3,1,2) - xc.grad).norm()/xc.grad.norm()).item(), " vs torch-channels_last:", ((xm.grad.permute(0,3,1,2) - xr.grad).norm()/xr.grad.norm()).item())
