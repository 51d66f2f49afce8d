"""World-size-2 gloo test (CPU) of the data-parallel gradient bucket: after launch()/finish() every
rank holds the mean of the per-rank gradients and p.grad aliases the flat bucket."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "multiple-objects-gan_b200")]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    from mog_b200 import parallel
    assert parallel.init_from_env("gloo") == world
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.BatchNorm1d(7), torch.nn.Linear(7, 3))
    if rank == 1:  # make rank 1 start different, then broadcast from 0
        for p in net.parameters():
            p.data.add_(1.0)
    parallel.broadcast_params(net)
    bucket = parallel.GradBucket(net.parameters())
    x = torch.full((4, 5), float(rank + 1))
    net(x).sum().backward()
    local = [p.grad.clone() for p in net.parameters()]
    bucket.launch()
    bucket.finish()
    gathered = [[torch.zeros_like(g) for _ in range(world)] for g in local]
    for g, lst in zip(local, gathered):
        dist.all_gather(lst, g)
    ok = True
    for p, lst in zip(net.parameters(), gathered):
        ok &= torch.allclose(p.grad, sum(lst) / world, atol=1e-6)
        ok &= p.grad.data_ptr() >= bucket.flat.data_ptr()
    # every view starts on a 16-byte boundary of the flat bucket (the fused optimiser's 128-bit path relies on it)
    for v in bucket.views:
        ok &= (v.data_ptr() - bucket.flat.data_ptr()) % 16 == 0
    # finish(scale=False): the bucket stays the SUM over ranks (1/world is folded into the fused Adam pass)
    net.zero_grad(set_to_none=True)
    net(x).sum().backward()
    local2 = [p.grad.clone() for p in net.parameters()]
    bucket.launch()
    bucket.finish(scale=False)
    for p, g in zip(net.parameters(), local2):
        lst = [torch.zeros_like(g) for _ in range(world)]
        dist.all_gather(lst, g)
        ok &= torch.allclose(p.grad, sum(lst), atol=1e-6)
    w0 = [p.detach().clone() for p in net.parameters()]
    ws = [[torch.zeros_like(w) for _ in range(world)] for w in w0]
    for w, lst in zip(w0, ws):
        dist.all_gather(lst, w)
        ok &= torch.equal(lst[0], lst[1])
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_grad_bucket_allreduce_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_single_rank_is_noop():
    sys.path[:0] = [os.path.join(ROOT, "multiple-objects-gan_b200")]
    from mog_b200 import parallel
    net = torch.nn.Linear(3, 2)
    net(torch.ones(1, 3)).sum().backward()
    g = net.weight.grad.clone()
    b = parallel.GradBucket(net.parameters())
    b.launch()
    b.finish()
    assert parallel.world() == 1 and torch.equal(net.weight.grad, g)
