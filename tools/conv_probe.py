"""Runs the dominant conv (G.h_net3.upsample: up2x + 3x3 96->96 @256^2, B=32) a few times: fwd, dgrad, wgrad.
Used under ncu --set full (profiles/) and for quick timing."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "multiple-objects-gan_b200")]
import torch
from mog_b200 import ops
from mog_b200._lib import PREC_NAMES
prec = PREC_NAMES[sys.argv[1] if len(sys.argv) > 1 else "bf16x3"]
which = sys.argv[2] if len(sys.argv) > 2 else "up"
B = 32
if which == "up":      # upBlock conv
    x = torch.randn(B, 128, 128, 96, device="cuda", requires_grad=True); w = (torch.randn(96, 96, 3, 3, device="cuda") * 0.03).requires_grad_(True); args = (1, 1, True)
elif which == "res":   # ResBlock conv 96->192 @128^2
    x = torch.randn(B, 128, 128, 96, device="cuda", requires_grad=True); w = (torch.randn(192, 96, 3, 3, device="cuda") * 0.03).requires_grad_(True); args = (1, 1, False)
elif which == "d256":  # D_NET256.img_code_s64: 1536->3072 4x4/s2 8->4
    x = torch.randn(B, 8, 8, 1536, device="cuda", requires_grad=True); w = (torch.randn(3072, 1536, 4, 4, device="cuda") * 0.01).requires_grad_(True); args = (2, 1, False)
for i in range(3):
    y = ops.conv2d(x, w, None, *args, 0, prec)
    y.backward(torch.ones_like(y))
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
ev[0].record(); y = ops.conv2d(x, w, None, *args, 0, prec); ev[1].record()
g = torch.ones_like(y); torch.cuda.synchronize()
ev[2].record(); y.backward(g); ev[3].record(); torch.cuda.synchronize()
print(which, "fwd %.3f ms  bwd(dgrad+wgrad) %.3f ms" % (ev[0].elapsed_time(ev[1]), ev[2].elapsed_time(ev[3])))
