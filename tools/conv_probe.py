"""Runs one of the dominant convs a few times: fwd, dgrad, wgrad, timed separately with CUDA events.
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
CASES = {
    "up": ((B, 128, 128, 96), (96, 96, 3, 3), (1, 1, True)),      # G.h_net3.upsample
    "res": ((B, 128, 128, 96), (192, 96, 3, 3), (1, 1, False)),   # ResBlock conv 96->192 @128^2
    "res2": ((B, 128, 128, 96), (96, 96, 3, 3), (1, 1, False)),   # ResBlock conv 96->96 @128^2
    "d2": ((B, 128, 128, 96), (192, 96, 4, 4), (2, 1, False)),    # D_NET256.img_code_s16.2: 4x4/s2 128->64
    "d256": ((B, 8, 8, 1536), (3072, 1536, 4, 4), (2, 1, False)), # D_NET256.img_code_s64: 8->4
    "img": ((B, 256, 256, 48), (3, 48, 3, 3), (1, 1, False)),     # GET_IMAGE_G @256^2
}
xs, ws, args = CASES[which]
x = torch.randn(*xs, device="cuda", requires_grad=True)
w = (torch.randn(*ws, device="cuda") * 0.03).requires_grad_(True)
for i in range(3):
    y = ops.conv2d(x, w, None, *args, 0, prec)
    y.backward(torch.ones_like(y))
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
ev[0].record(); y = ops.conv2d(x, w, None, *args, 0, prec); ev[1].record()
g = torch.ones_like(y); torch.cuda.synchronize()
ev[2].record(); (gx,) = torch.autograd.grad(y, x, g, retain_graph=True); ev[3].record(); torch.cuda.synchronize()
ev[4].record(); (gw,) = torch.autograd.grad(y, w, g); ev[5].record(); torch.cuda.synchronize()
print(which, "fwd %.3f ms  dgrad(+split) %.3f ms  wgrad(+split) %.3f ms" % (ev[0].elapsed_time(ev[1]), ev[2].elapsed_time(ev[3]), ev[4].elapsed_time(ev[5])))
