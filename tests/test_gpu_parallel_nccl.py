"""Two ranks over NCCL against one rank on the concatenated batch, where the semantics allow the comparison:
a BatchNorm-free libmog conv stack -- per-rank batch statistics are what differs between data-parallel replicas and one big
batch, everything else must agree.  Each rank runs forward / backward on its half of the batch through the product path
(bf16x3 convs, ``GradBucket`` all-reduce, fused Adam with the 1/world scale folded in); rank 0 also runs the single-rank step
on the whole batch and compares the parameters after two optimiser steps.

Needs two GPUs: skipped on a one-GPU box (run with ``gpurun --gpus 2``).  Run on the B200 box: -m gpu."""
import os
import socket
import sys

import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.gpu

# the two-rank sum of half-batch gradients and the one-rank full-batch gradient differ by fp32 summation order only
TOL = 2e-5


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _make(device):
    g = torch.Generator().manual_seed(3)
    shapes = [(32, 8, 3, 3), (48, 32, 4, 4), (3, 48, 3, 3)]
    return [(torch.randn(s, generator=g) * (1.0 / (s[1] * s[2] * s[3]) ** 0.5)).to(device).requires_grad_(True) for s in shapes]


def _loss(ws, x):
    from mog_b200 import _lib, ops
    h = ops.conv2d(x, ws[0], None, 1, 1, False, _lib.ACT_LRELU)
    h = ops.conv2d(h, ws[1], None, 2, 1, False, _lib.ACT_LRELU)
    y = ops.conv2d(h, ws[2], None, 1, 1, False, _lib.ACT_TANH)
    return y.square().mean()


def _worker(rank, world, port, q):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "multiple-objects-gan_b200")]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch.distributed as dist
    from mog_b200 import ops, parallel
    from mog_b200.optim import Adam
    try:
        assert parallel.init_from_env("nccl") == world
        dev = torch.device("cuda", rank)
        ops.set_precision("bf16x3")
        g = torch.Generator().manual_seed(7)
        xs = [torch.randn(8, 16, 16, 8, generator=g).to(dev) for _ in range(2)]     # the global batch of each step
        half = 8 // world
        ws = _make(dev)
        for w in ws:                          # (same seed on both ranks; broadcast as the trainers do)
            dist.broadcast(w.data, 0)
        ops.invalidate_packed(ws)
        bucket = parallel.GradBucket(ws)
        opt = Adam(ws, lr=1e-2, betas=(0.5, 0.999))
        for x in xs:
            opt.zero_grad(set_to_none=True)
            _loss(ws, x[rank * half:(rank + 1) * half]).backward()
            bucket.launch()
            bucket.finish(scale=False)
            opt.step(grad_scale=1.0 / world)
        torch.cuda.synchronize()
        ok, worst = True, 0.0
        if rank == 0:
            ref = _make(dev)
            opt1 = Adam(ref, lr=1e-2, betas=(0.5, 0.999))
            for x in xs:
                opt1.zero_grad(set_to_none=True)
                _loss(ref, x).backward()          # mean over the whole batch = mean of the two half-batch means
                opt1.step()
            for a, b in zip(ws, ref):
                e = float((a.detach().double() - b.detach().double()).norm() / b.detach().double().norm())
                worst = max(worst, e)
            ok = worst < TOL
        # both ranks hold the same parameters
        for w in ws:
            lst = [torch.zeros_like(w) for _ in range(world)]
            dist.all_gather(lst, w.detach())
            ok = ok and torch.equal(lst[0], lst[1])
        q.put((rank, bool(ok), worst))
        dist.barrier()
    except Exception as e:  # noqa: BLE001
        q.put((rank, False, repr(e)[:300]))
    os._exit(0)      # (no NCCL teardown: see bench.py)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_two_ranks_equal_one_rank_on_the_concatenated_batch():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=240) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert [r[:2] for r in res] == [(0, True), (1, True)], res
