import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "multiple-objects-gan_b200")]
import torch, torch.nn as nn
from mog_b200 import ops
from mog_b200.ops import ACT_GLU
ops.set_precision("bf16x3")
M, C = 32 * 16384, 192
x = torch.randn(32, 128, 128, C, device="cuda", requires_grad=True)
bn = nn.BatchNorm2d(C).cuda().train()
for _ in range(2):
    y = ops.bn_act(x, bn, ACT_GLU); y.backward(torch.randn_like(y)); x.grad = None
h = torch.randn(32, 16384, 48, device="cuda", requires_grad=True); src = torch.randn(32, 18, 48, device="cuda", requires_grad=True)
mask = torch.zeros(32, 18, dtype=torch.bool, device="cuda"); mask[:, 12:] = True
for _ in range(2):
    out, attn = ops.word_attention(h, src, mask); out.backward(torch.randn_like(out)); h.grad = None; src.grad = None
torch.cuda.synchronize()
