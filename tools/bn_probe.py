"""Times the BatchNorm kernels alone (CUDA events, L2 flushed) at the generator's big shapes: GB/s vs algorithmic bytes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "multiple-objects-gan_b200")]
import torch
from mog_b200._lib import call
from mog_b200 import ops
st = lambda: torch.cuda.current_stream().cuda_stream
flush = torch.empty(64 * 1024 * 1024, device="cuda")
def timeit(fn, n=5):
    ts = []
    for _ in range(n):
        flush.zero_(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        fn(); a.record(); fn(); fn(); fn(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b) / 4)   # back to back: no launch gaps
    return sorted(ts)[len(ts) // 2]
for (M, C, act) in [(32 * 65536, 96, 3), (32 * 16384, 192, 3), (32 * 16384, 96, 0), (32 * 4096, 192, 3), (2 * 32 * 1024, 384, 2)]:
    Co = C // 2 if act == 3 else C
    x = torch.randn(M, C, device="cuda"); dy = torch.randn(M, Co, device="cuda")
    from mog_b200._lib import lib
    P0, P1 = lib().mog_bn_parts(1, M, C, act, 0), lib().mog_bn_parts(1, M, C, act, 1)
    part = torch.zeros(P0, 2, 1, C, device="cuda", dtype=torch.float64); mis = torch.zeros(4, 1, C, device="cuda")
    g = torch.ones(C, device="cuda"); b = torch.zeros(C, device="cuda"); y = torch.empty(M, Co, device="cuda")
    red = torch.zeros(P1 + 1, 2, 1, C, device="cuda", dtype=torch.float64); dx = torch.empty_like(x); dgb = torch.empty(2, C, device="cuda")
    f_stats = lambda: call("mog_bn_stats", x.data_ptr(), 1, M, C, part.data_ptr(), P0, st())
    f_stats()
    call("mog_bn_finalize", part.data_ptr(), P0, 1, M, C, g.data_ptr(), b.data_ptr(), 1e-5, 0.1, None, None,
         mis[0].data_ptr(), mis[1].data_ptr(), mis[2].data_ptr(), mis[3].data_ptr(), st())
    f_fwd = lambda: call("mog_affine_act_fwd", x.data_ptr(), mis[2].data_ptr(), mis[3].data_ptr(), None, y.data_ptr(), 1, M, C, act, st())
    f_red = lambda: call("mog_bn_act_bwd_reduce", x.data_ptr(), dy.data_ptr(), mis[0].data_ptr(), mis[1].data_ptr(), g.data_ptr(), b.data_ptr(), 1, M, C, act, red.data_ptr(), P1, red[P1, 0].data_ptr(), red[P1, 1].data_ptr(), dgb[0].data_ptr(), dgb[1].data_ptr(), st())
    f_app = lambda: call("mog_bn_act_bwd_apply", x.data_ptr(), dy.data_ptr(), mis[0].data_ptr(), mis[1].data_ptr(), g.data_ptr(), b.data_ptr(), red[P1, 0].data_ptr(), red[P1, 1].data_ptr(), 1, M, C, act, dx.data_ptr(), st())
    for name, f, nbytes in (("stats", f_stats, 4 * M * C), ("fwd", f_fwd, 4 * M * (C + Co)), ("bwd_reduce", f_red, 4 * M * (C + Co)), ("bwd_apply", f_app, 4 * M * (2 * C + Co))):
        f(); ms = timeit(f)
        print("M=%8d C=%3d act=%d %-10s %7.1f us %6.0f GB/s" % (M, C, act, name, ms * 1e3, nbytes / ms / 1e6))
