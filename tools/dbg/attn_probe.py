import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "multiple-objects-gan_b200")]
import torch
from mog_b200 import ops
flush = torch.empty(64 * 1024 * 1024, device="cuda")
def timeit(fn, n=5):
    ts = []
    for _ in range(n):
        flush.zero_(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        fn(); a.record(); fn(); fn(); fn(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b) / 4)
    return sorted(ts)[len(ts) // 2]
for (B, Q, D, T) in [(32, 16384, 48, 18), (32, 4096, 48, 18)]:
    h = torch.randn(B, Q, D, device="cuda", requires_grad=True); src = torch.randn(B, T, D, device="cuda", requires_grad=True)
    mask = torch.zeros(B, T, dtype=torch.bool, device="cuda"); mask[:, 12:] = True
    out, attn = ops.word_attention(h, src, mask); g = torch.randn_like(out)
    f = lambda: ops.word_attention(h, src, mask)
    ms = timeit(f); print("attn fwd B%d Q%d: %.1f us  %.0f GB/s" % (B, Q, ms * 1e3, (2 * 4 * B * Q * D + 4 * B * T * Q) / ms / 1e6))
    def bw():
        h.grad = None; src.grad = None
        out.backward(g, retain_graph=True)
    ms = timeit(bw); print("attn bwd B%d Q%d: %.1f us  %.0f GB/s" % (B, Q, ms * 1e3, (3 * 4 * B * Q * D) / ms / 1e6))
