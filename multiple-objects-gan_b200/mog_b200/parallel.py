"""Data parallelism for the G/D step: one process per GPU, parameters resident, gradients summed
with NCCL all-reduce over NVLink/NVSwitch (``torch.distributed``), nothing else crosses GPUs.

The reference uses single-process ``nn.parallel.data_parallel`` (re-broadcasting all parameters
on each of its 10 calls per step and gathering activations on GPU 0; trainer.py:296,
losses.py:146-152,193).  Here each rank runs the whole step on its own shard of the batch --
BatchNorm statistics are per shard exactly as in the reference's replicas -- and each network's
gradients are all-reduced as ONE flat bucket right after its backward, asynchronously, so the
643 MB of D_NET256 travel while the next network computes.

Semantic differences vs. the reference's DataParallel are documented in DESIGN.md (loss heads /
wrong-pair shift / BN of the heads are per shard here, global on GPU 0 there).
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def world():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def rank():
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def init_from_env(backend=None):
    """Initialise torch.distributed from torchrun's RANK/WORLD_SIZE/MASTER_* (no-op for 1 rank)."""
    ws = int(os.environ.get("WORLD_SIZE", "1"))
    if ws <= 1 or dist.is_initialized():
        return ws
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group(backend=backend)
    return ws


class GradBucket:
    """Flat fp32 gradient bucket of one network.  ``launch()`` copies the .grad tensors into the
    flat buffer and starts an asynchronous all-reduce(sum); ``finish()`` waits, scales by
    1/world (mean over the global batch) and makes every ``p.grad`` a view into the bucket."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        # every view starts on a 16-byte boundary so the fused optimiser keeps its 128-bit path
        n = sum((p.numel() + 3) // 4 * 4 for p in self.params)
        dev = self.params[0].device if self.params else "cpu"
        self.flat = torch.zeros(n, device=dev, dtype=torch.float32)
        self.views = []
        off = 0
        for p in self.params:
            self.views.append(self.flat[off:off + p.numel()].view_as(p))
            off += (p.numel() + 3) // 4 * 4
        self.handle = None

    def launch(self):
        if world() == 1:
            return
        src, dst = [], []
        for p, v in zip(self.params, self.views):
            if p.grad is None:
                v.zero_()
            elif p.grad.data_ptr() != v.data_ptr():
                src.append(p.grad)
                dst.append(v)
        if dst:
            torch._foreach_copy_(dst, src)     # one multi-tensor launch instead of one copy per parameter
        self.handle = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, async_op=True)

    def finish(self, scale=True):
        """``scale=False`` leaves the bucket as the SUM over ranks (the caller folds 1/world into the optimiser:
        ``mog_b200.optim.Adam.step(grad_scale=1/world)``)."""
        if world() == 1:
            return
        if self.handle is not None:
            self.handle.wait()
            self.handle = None
        if scale:
            self.flat.mul_(1.0 / world())
        for p, v in zip(self.params, self.views):
            p.grad = v


def broadcast_params(module, src=0):
    """Make every rank start from rank ``src``'s parameters and buffers."""
    if world() == 1:
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src)
    from . import ops
    ops.invalidate_packed(module)
