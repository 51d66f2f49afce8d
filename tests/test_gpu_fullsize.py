"""Properties at the sizes bench.py runs (BASELINE.json config 5, B = 32 per GPU) -- where the CPU oracle cannot go
(one oracle step at B = 32 is minutes).  Size-independent properties of the path instead of element-wise references:

* adjointness of every convolution form on the benchmark's own layer shapes: y = conv(x, w) is bilinear, so for any g
  <g, conv(x, w)> = <dgrad(g), x> = <wgrad(g), w>.  Forward, data gradient and weight gradient run on different kernels
  (halo / small-grid / patch-matrix forms, split-K, sub-pixel phases, parity views), so the three inner products agreeing pins
  them against each other at full size; the forward itself is pinned element-wise at small sizes (test_gpu_tc.py) and at
  config-5 widths with B = 4 (test_gpu_config5.py).
* the full training step is bit-reproducible run to run, and its CUDA-graph form (what bench.py times) leaves bit-identical
  parameters, EMA copy, Adam moments and BatchNorm buffers as the eager step -- with the independent branches of the step on
  separate CUDA streams (cfg.MOG.STREAMS, the default) and on one stream.

Run on the B200 box: -m gpu."""
import numpy as np
import pytest
import torch

from mog_b200 import synth

pytestmark = pytest.mark.gpu

# bf16x3 keeps ~16 mantissa bits per product; the inner products are sums of 1e7..1e9 such terms with random signs.
ADJ_TOL = 2e-4

FULL_SIZE_LAYERS = [
    # N, H, W, Cin, Cout, KH, KW, stride, pad, up2x     (shapes of profiles/r2z_shape_profile.txt)
    (32, 128, 128, 96, 192, 3, 3, 1, 1, False),   # ResBlock conv C -> 2C @128^2 (halo form)
    (32, 128, 128, 96, 96, 3, 3, 1, 1, True),     # G.h_net3.upsample: up2x + 3x3 -> 256^2 (sub-pixel phases) -- the roofline conv
    (64, 128, 128, 96, 192, 4, 4, 2, 1, False),   # D_NET256 layer 2, real+fake pair pass (parity views)
    (64, 8, 8, 1536, 3072, 4, 4, 2, 1, False),    # D_NET256.img_code_s64 (small-grid form, split-K, 302 MB weight gradient)
    (64, 4, 4, 3072, 1536, 3, 3, 1, 1, False),    # D_NET256.img_code_s64_1 (3x3 on 4x4)
    (64, 256, 256, 3, 96, 4, 4, 2, 1, False),     # D_NET256 layer 1: 3 input channels (patch-matrix forward / weight gradient)
    (32, 256, 256, 48, 3, 3, 3, 1, 1, False),     # GET_IMAGE_G @256^2: 3 output channels (patch-matrix backward)
    (32, 17, 17, 768, 192, 1, 1, 1, 0, False),    # Inception 17x17 1x1 (re-tiled pixel grid, single-sub-tile items)
    (32, 17, 17, 192, 192, 7, 1, 1, (3, 0), False),   # Inception 7x1 (frozen encoder: forward + data gradient only)
    (32, 17, 17, 192, 192, 1, 7, 1, (0, 3), False),   # Inception 1x7 (rows of all images stacked into one image)
    (32, 35, 35, 48, 64, 5, 5, 1, 2, False),      # Inception 5x5
    (32, 299, 299, 3, 32, 3, 3, 2, 0, False),     # Inception stem: 3 input channels, 3x3/s2 without padding, odd size
]


def _dot(a, b):
    return float((a.double() * b.double()).sum())


def _id(c):
    return "x".join("%d-%d" % v if isinstance(v, tuple) else str(int(v)) for v in c)


@pytest.mark.parametrize("case", FULL_SIZE_LAYERS, ids=_id)
def test_conv_adjointness_at_benchmark_shapes(case):
    from mog_b200 import ops
    N, H, W, Ci, Co, KH, KW, s, p, up = case
    frozen = isinstance(p, tuple)     # (different padding per axis: the weight gradient is not implemented -- nor needed)
    ops.set_precision("bf16x3")
    try:
        g0 = torch.Generator(device="cuda").manual_seed(11)
        x = torch.randn(N, H, W, Ci, device="cuda", generator=g0).requires_grad_(True)
        w = (torch.randn(Co, Ci, KH, KW, device="cuda", generator=g0) / np.sqrt(Ci * KH * KW)).requires_grad_(not frozen)
        pad = list(p) if frozen else p
        y = ops.conv2d(x, w, None, s, pad, up, 0)
        g = torch.randn(y.shape, device="cuda", generator=g0)
        grads = torch.autograd.grad(y, (x,) if frozen else (x, w), g)
        torch.cuda.synchronize()
        assert torch.isfinite(y).all() and all(torch.isfinite(t).all() for t in grads)
        fy, fx = _dot(g, y.detach()), _dot(grads[0], x.detach())
        # the inner product of independent Gaussians has zero mean: normalise by ||g|| ||y||
        scale = float(g.double().norm() * y.detach().double().norm())
        assert abs(fy - fx) <= ADJ_TOL * scale, ("dgrad", fy, fx, scale)
        if not frozen:
            fw = _dot(grads[1], w.detach())
            assert abs(fy - fw) <= ADJ_TOL * scale, ("wgrad", fy, fw, scale)
            # the weight gradient of the same inputs is bit-reproducible (fixed-order split reductions)
            (dw2,) = torch.autograd.grad(ops.conv2d(x, w, None, s, pad, up, 0), (w,), g)
            assert torch.equal(grads[1], dw2)
    finally:
        ops.set_precision("fp32")


def _snapshot(st):
    out = []
    for net in [st["netG"]] + st["netsD"]:
        out += [p.detach().clone() for p in net.parameters()] + [b.detach().clone() for b in net.buffers()]
    out += [a.detach().clone() for a in st["avg_param_G"]]
    for opt in [st["optG"]] + st["optDs"]:
        for s in opt.state_dict()["state"].values():
            out += [s["exp_avg"].detach().clone(), s["exp_avg_sq"].detach().clone()]
    return out


def test_full_size_step_is_reproducible_and_graph_equals_eager():
    """Config 5 of BASELINE.json at B = 32 with the DAMSM branch through the libmog Inception-v3 -- the step bench.py times."""
    import bench
    from mog_b200 import ops
    from mog_b200.attngan.model import CNN_ENCODER
    from mog_b200.attngan.trainer import condGANTrainer
    B, K = 32, 3
    cfg = bench.set_cfg()
    cfg.TRAIN.BATCH_SIZE = B
    cfg.MOG.PRECISION = "bf16x3"
    h = synth.attngan_batch(B, seed=1234)
    d = {k: v.cuda() for k, v in h.items() if torch.is_tensor(v)}
    imgs = [t.cuda() for t in h["imgs"]]
    args = (imgs, d["sent_emb"], d["words_embs"], d["mask"], d["transf_matrices"], d["transf_matrices_inv"], d["label_one_hot"],
            h["cap_lens"], h["class_ids"])
    rng = np.random.RandomState(3)
    noises = [torch.from_numpy(rng.standard_normal((B, cfg.GAN.Z_DIM)).astype(np.float32)).cuda() for _ in range(K)]
    epss = [torch.from_numpy(rng.standard_normal((B, 100)).astype(np.float32)).cuda() for _ in range(K)]

    def make():
        torch.manual_seed(1234)
        tr = condGANTrainer("", None, 0, None)
        enc = CNN_ENCODER(256)
        enc.load_state_dict(synth.fill_encoder_state_dict(enc.state_dict(), 9))
        for p in enc.parameters():
            p.requires_grad = False
        enc.cuda().eval()
        _, _, netG, netsD, _ = tr.build_models(image_encoder=enc, load_encoders=False)
        optG, optDs = tr.define_optimizers(netG, netsD)
        return tr, tr.make_step_state(netG, netsD, optG, optDs)

    try:
        runs = []
        for mode in ("eager", "eager", "graph", "one-stream"):
            cfg.MOG.STREAMS = mode != "one-stream"     # default: the independent branches of the step on separate streams
            tr, st = make()
            losses = []
            if mode in ("eager", "one-stream"):
                for k in range(K):
                    losses.append([float(t) for t in tr.train_step(st, *args, noise=noises[k], eps=epss[k])])
            else:   # one eager step (creates the optimiser state), capture without training, then replays
                losses.append([float(t) for t in tr.train_step(st, *args, noise=noises[0], eps=epss[0])])
                gs = tr.graphed_step(st, *args, warmup=1, noise=noises[1], eps=epss[1], dry_warmup=True)
                for k in range(1, K):
                    losses.append([float(t) for t in gs(*args, noise=noises[k], eps=epss[k])])
            torch.cuda.synchronize()
            runs.append((losses, _snapshot(st)))
            del tr, st
            torch.cuda.empty_cache()
        (l0, s0), (l1, s1), (l2, s2), (l3, s3) = runs
        assert all(np.isfinite(v) for step in l0 for v in step)
        assert l0 == l1, "two eager runs of the same step differ"
        assert all(torch.equal(a, b) for a, b in zip(s0, s1)), "the eager step is not bit-reproducible"
        assert l0 == l2, ("graph vs eager losses", l0, l2)
        bad = [i for i, (a, b) in enumerate(zip(s0, s2)) if not torch.equal(a, b)]
        assert not bad, "graph replay differs from the eager step in %d of %d state tensors" % (len(bad), len(s0))
        assert l0 == l3 and all(torch.equal(a, b) for a, b in zip(s0, s3)), "multi-stream step differs from the single-stream step"
    finally:
        cfg.MOG.STREAMS = True
        ops.set_precision("fp32")
