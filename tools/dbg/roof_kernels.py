"""Launches the roofline conv (G.h_net3.upsample fwd / dgrad / wgrad) and a deep discriminator wgrad once each (for ncu --set full)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "multiple-objects-gan_b200")]
import torch
from mog_b200 import ops
ops.set_precision("bf16x3")
x = torch.randn(32, 128, 128, 96, device="cuda", requires_grad=True)
w = (torch.randn(96, 96, 3, 3, device="cuda") * 0.03).requires_grad_(True)
for _ in range(2):
    y = ops.conv2d(x, w, None, 1, 1, True, 0); y.backward(torch.randn_like(y)); x.grad = None; w.grad = None
x2 = torch.randn(64, 8, 8, 1536, device="cuda", requires_grad=True)
w2 = (torch.randn(3072, 1536, 4, 4, device="cuda") * 0.01).requires_grad_(True)
for _ in range(2):
    y2 = ops.conv2d(x2, w2, None, 2, 1, False, 0); y2.backward(torch.randn_like(y2)); x2.grad = None; w2.grad = None
torch.cuda.synchronize()
