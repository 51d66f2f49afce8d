"""Debug helper: run the tcgen05 conv cases one by one and print rel-L2 errors (no asserts)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "multiple-objects-gan_b200"), os.path.join(ROOT, "tests")]
import numpy as np, torch
import test_gpu_tc as T
from mog_b200 import ops
from mog_b200._lib import PREC_NAMES

precs = sys.argv[1].split(",") if len(sys.argv) > 1 else ["bf16x3"]
for prec in precs:
    for case in (T.TMA_CASES if 'tma' in sys.argv else T.TC_CASES):
        N, H, W, Ci, Co, k, s, p, up, has_b, act = case
        x = T.rnd(N, Ci, H, W, seed=1).requires_grad_(True)
        w = T.rnd(Co, Ci, k, k, seed=2, scale=1.0 / np.sqrt(Ci * k * k)).requires_grad_(True)
        b = T.rnd(Co, seed=3, scale=0.1).requires_grad_(True) if has_b else None
        y_ref = T._torch_conv(x, w, b, s, p, up, act)
        g = T.rnd(*y_ref.shape, seed=4)
        y_ref.backward(g)
        xd = x.detach().cuda().permute(0, 2, 3, 1).contiguous().requires_grad_(True)
        wd = w.detach().cuda().requires_grad_(True)
        bd = b.detach().cuda().requires_grad_(True) if has_b else None
        try:
            y = ops.conv2d(xd, wd, bd, s, p, up, act, precision=PREC_NAMES[prec])
            torch.cuda.synchronize()
            ef = T.rel(y.permute(0, 3, 1, 2), y_ref)
            y.backward(g.cuda().permute(0, 2, 3, 1).contiguous())
            torch.cuda.synchronize()
            ed = T.rel(xd.grad.permute(0, 3, 1, 2), x.grad)
            ew = T.rel(wd.grad, w.grad)
            print("%-7s %-40s fwd %.2e dgrad %.2e wgrad %.2e" % (prec, case, ef, ed, ew), flush=True)
        except Exception as e:
            print("%-7s %-40s EXC %s" % (prec, case, e), flush=True)
            sys.exit(1)
