"""DAMSM image encoder (CNN_ENCODER = frozen Inception-v3 trunk + two projections): oracle pinned against the executed
reference (CPU), state_dict contract (CPU), libmog parity through the C ABI (GPU), pooling / resize kernels vs torch (GPU)."""
import json
import os

import pytest
import torch
import torch.nn.functional as F

import golden_util as gu
from mog_b200 import synth
from oracle import encoder_oracle as EO

HERE = os.path.dirname(os.path.abspath(__file__))


def _full_state(meta):
    """Same deterministic weights as tests/golden/make_golden_encoder.py (which fills torchvision's full inception_v3
    dict, incl. the AuxLogits / fc entries the encoder drops, then the emb_* entries with seed + 1)."""
    shapes = json.load(open(os.path.join(HERE, "golden", "inception_full_shapes.json")))
    full = synth.fill_encoder_state_dict({k: torch.empty(s) for k, s in shapes.items()}, meta["seed"])
    keys = json.load(open(os.path.join(HERE, "golden", "cnn_encoder_keys.json")))
    own = synth.fill_encoder_state_dict({k: torch.empty(s) for k, s in keys.items() if k.startswith("emb_")}, meta["seed"] + 1)
    return {k: (own[k] if k.startswith("emb_") else full[k]) for k in keys}, keys


def test_oracle_matches_reference():
    G, meta = gu.load("cnn_encoder")
    P, _ = _full_state(meta)
    img, pf, pc = synth.encoder_probe(meta["B"], meta["nef"], meta["seed"] + 2)
    img.requires_grad_(True)
    feat, code = EO.cnn_encoder(P, img)
    gu.check(feat, G["features"], 2e-5, "features")
    gu.check(code, G["cnn_code"], 2e-5, "cnn_code")
    loss = (feat * pf).sum() + (code * pc).sum()
    gu.check(loss, G["loss"], 2e-5, "loss")
    (g,) = torch.autograd.grad(loss, img)
    gu.check(g, G["d_img"], 1e-4, "d_img")


def test_state_dict_contract():
    from mog_b200.attngan import model as M
    from mog_b200.attngan.miscc.config import cfg, reset_cfg
    reset_cfg()
    _, meta = gu.load("cnn_encoder")
    keys = json.load(open(os.path.join(HERE, "golden", "cnn_encoder_keys.json")))
    enc = M.CNN_ENCODER(meta["nef"])
    sd = {k: list(v.shape) for k, v in enc.state_dict().items()}
    assert sd == keys and list(sd) == list(keys)
    assert not any(p.requires_grad for n, p in enc.named_parameters() if not n.startswith("emb_"))


@pytest.mark.gpu
@pytest.mark.parametrize("prec", ["fp32", "bf16x3"])
def test_libmog_encoder_matches_reference(prec):
    from mog_b200 import ops
    from mog_b200.attngan import model as M
    from mog_b200.attngan.miscc.config import reset_cfg
    reset_cfg()
    G, meta = gu.load("cnn_encoder")
    P, _ = _full_state(meta)
    ops.set_precision(prec)
    try:
        enc = M.CNN_ENCODER(meta["nef"])
        enc.load_state_dict(P)
        for p in enc.parameters():
            p.requires_grad = False
        enc.cuda().eval()
        img, pf, pc = (t.cuda() for t in synth.encoder_probe(meta["B"], meta["nef"], meta["seed"] + 2))
        img.requires_grad_(True)
        feat, code = enc(img)
        tol_o, tol_g = (5e-5, 2e-2) if prec == "fp32" else (2e-4, 3e-2)   # gradient: ReLU mask flips through ~95 layers
        gu.check(feat, G["features"], tol_o, "features")
        gu.check(code, G["cnn_code"], tol_o, "cnn_code")
        loss = (feat * pf).sum() + (code * pc).sum()
        loss.backward()
        gu.check(loss, G["loss"], tol_o, "loss")
        gu.check(img.grad, G["d_img"], tol_g, "d_img")
        with pytest.raises(RuntimeError):
            enc.train()(img)
    finally:
        ops.set_precision("fp32")


@pytest.mark.gpu
@pytest.mark.parametrize("case", [(0, 3, 2, 0, 35, 35, 24), (0, 3, 2, 0, 147, 147, 24), (1, 3, 1, 1, 17, 17, 24), (1, 8, 8, 0, 8, 8, 24),
                                  (0, 3, 2, 1, 9, 12, 24), (1, 2, 2, 0, 7, 9, 24), (0, 3, 2, 0, 13, 11, 6), (1, 3, 1, 1, 5, 7, 3),
                                  # wider channel counts (several 128-byte lines per pixel)
                                  (0, 3, 2, 0, 35, 35, 64), (1, 3, 1, 1, 17, 17, 96), (0, 3, 2, 1, 9, 12, 32), (1, 2, 2, 0, 7, 9, 64),
                                  (0, 3, 1, 1, 6, 21, 32)])
def test_pool2d_matches_torch(case):
    from mog_b200 import ops
    mode, k, s, p, H, W, Cc = case
    torch.manual_seed(3)
    x = torch.randn(3, H, W, Cc, device="cuda")
    x[0, :4, :4] = 1.25   # ties: the first maximum must get the gradient
    # (contiguous NCHW on the torch side: its avg_pool2d backward mishandles a channels_last-strided gradient on this build)
    xr = x.permute(0, 3, 1, 2).contiguous().requires_grad_(True)
    xm = x.clone().requires_grad_(True)
    fn = ops.max_pool2d if mode == 0 else ops.avg_pool2d
    y = fn(xm, k, s, p)
    yr = (F.max_pool2d if mode == 0 else F.avg_pool2d)(xr, k, s, p)
    assert torch.allclose(y.permute(0, 3, 1, 2), yr, rtol=1e-6, atol=1e-6)
    g = torch.randn_like(y)
    y.backward(g)
    yr.backward(g.permute(0, 3, 1, 2).contiguous())
    assert torch.allclose(xm.grad.permute(0, 3, 1, 2), xr.grad, rtol=1e-5, atol=2e-6)
    if mode == 0:   # the C-ABI's other max-pool backward: arg-max recomputed from the forward input (no recorded positions)
        from mog_b200._lib import call
        dx = torch.empty_like(x)
        N, Ho, Wo = y.shape[0], y.shape[1], y.shape[2]
        call("mog_pool2d_bwd", x.data_ptr(), None, g.contiguous().data_ptr(), dx.data_ptr(), N, H, W, Cc, k, s, p, 0,
             torch.cuda.current_stream().cuda_stream)
        assert torch.allclose(dx.permute(0, 3, 1, 2), xr.grad, rtol=1e-5, atol=2e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("align", [False, True])
@pytest.mark.parametrize("size", [(256, 256, 299, 299), (17, 23, 40, 31), (8, 8, 8, 8)])
def test_resize_bilinear_matches_torch(size, align):
    from mog_b200 import ops
    Hi, Wi, Ho, Wo = size
    torch.manual_seed(4)
    x = torch.randn(2, Hi, Wi, 3, device="cuda")
    xr = x.permute(0, 3, 1, 2).contiguous().requires_grad_(True)
    xm = x.clone().requires_grad_(True)
    y = ops.resize_bilinear(xm, (Ho, Wo), align)
    yr = F.interpolate(xr, size=(Ho, Wo), mode="bilinear", align_corners=align)
    assert torch.allclose(y.permute(0, 3, 1, 2), yr, rtol=1e-5, atol=1e-5)
    g = torch.randn_like(y)
    y.backward(g)
    yr.backward(g.permute(0, 3, 1, 2).contiguous())
    assert torch.allclose(xm.grad.permute(0, 3, 1, 2), xr.grad, rtol=1e-4, atol=1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("prec", ["fp32", "bf16x3"])
@pytest.mark.parametrize("geom", [((1, 7), (0, 3), 1, 17, 17, 128, 192), ((7, 1), (3, 0), 1, 17, 17, 160, 160),
                                  ((1, 3), (0, 1), 1, 8, 8, 384, 384), ((3, 1), (1, 0), 1, 8, 8, 384, 384),
                                  ((5, 5), (2, 2), 1, 35, 35, 48, 64), ((3, 3), (0, 0), 2, 35, 35, 288, 384),
                                  ((3, 3), (0, 0), 2, 299, 299, 3, 32), ((3, 3), (0, 0), 1, 73, 73, 80, 192)])
def test_rect_conv_fwd_dgrad_matches_torch(geom, prec):
    """Convolutions with different padding / filter extent along H and W (forward + data gradient, frozen weight).
    The data gradient is compared without the activation: a ReLU mask computed from outputs that agree to 5e-6 still
    flips for the handful of pre-activations within 1e-6 of zero, and each flip is a full-size error in three rows."""
    from mog_b200 import ops
    ks, pad, stride, H, W, Ci, Co = geom
    torch.backends.cudnn.allow_tf32 = False      # the torch side is the fp32 reference here
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(5)
    N = 2
    x = torch.randn(N, H, W, Ci, device="cuda")
    w = torch.randn(Co, Ci, ks[0], ks[1], device="cuda") / (Ci * ks[0] * ks[1]) ** 0.5
    b = torch.randn(Co, device="cuda") * 0.1
    xr = x.permute(0, 3, 1, 2).detach().clone().requires_grad_(True)
    xm = x.clone().requires_grad_(True)
    mpad = pad if pad[0] != pad[1] else pad[0]
    tol = 2e-5 if prec == "fp32" else 5e-5
    with torch.no_grad():
        ya = ops.conv2d(x, w, b, stride, mpad, False, ops.ACT_RELU, ops.PREC_NAMES[prec])
        assert gu.rel_l2(ya.permute(0, 3, 1, 2).cpu().numpy(), F.relu(F.conv2d(xr, w, b, stride, pad)).cpu().numpy()) < tol
    y = ops.conv2d(xm, w, b, stride, mpad, False, ops.ACT_NONE, ops.PREC_NAMES[prec])
    yr = F.conv2d(xr, w, b, stride, pad)
    assert gu.rel_l2(y.permute(0, 3, 1, 2).detach().cpu().numpy(), yr.detach().cpu().numpy()) < tol
    g = torch.randn_like(y)
    y.backward(g)
    yr.backward(g.permute(0, 3, 1, 2))
    assert gu.rel_l2(xm.grad.permute(0, 3, 1, 2).cpu().numpy(), xr.grad.cpu().numpy()) < tol
