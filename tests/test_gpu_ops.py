"""Per-kernel parity: libmog (through the C ABI / ops wrappers) vs. the oracle primitives
(torch CPU fp32) and the reference-generated golden vectors.  Run on the B200 box: -m gpu."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import golden_util as gu
from oracle import attngan_oracle as O

pytestmark = pytest.mark.gpu

FP32_TOL = 2e-5   # fp32 kernels vs fp32 CPU: only summation order differs


def _ops():
    from mog_b200 import ops
    return ops


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g) * scale


CONV_CASES = [
    # N, H, W, Cin, Cout, k, stride, pad, up2x, bias, act
    (2, 8, 8, 16, 32, 3, 1, 1, False, False, 0),
    (2, 8, 8, 16, 32, 3, 1, 1, True, False, 0),        # upBlock: fused nearest x2
    (3, 16, 16, 96, 192, 3, 1, 1, False, False, 0),    # ResBlock conv (G hot shape, small spatial)
    (2, 16, 16, 3, 96, 4, 2, 1, False, False, 2),      # D first conv + LeakyReLU epilogue (Cin=3 scalar path)
    (2, 16, 16, 96, 192, 4, 2, 1, False, False, 0),    # downBlock
    (3, 16, 16, 84, 40, 4, 1, 1, False, False, 0),     # D_NET64.local: 16 -> 15
    (2, 16, 16, 100, 50, 3, 2, 1, False, False, 2),    # bbox_net stride-2 3x3
    (2, 9, 9, 25, 12, 3, 2, 1, False, False, 0),       # odd sizes / scalar channels
    (4, 4, 4, 64, 1, 4, 4, 0, False, True, 5),         # outlogits: 4x4/s4 + bias + sigmoid
    (2, 32, 32, 48, 3, 3, 1, 1, False, False, 4),      # GET_IMAGE_G: Cout=3 + tanh
    (5, 1, 1, 181, 100, 1, 1, 0, False, False, 0),     # Linear 181 -> 100
    (4, 1, 1, 32, 400, 1, 1, 0, False, True, 0),       # Linear + bias (CA_NET)
    (2, 6, 1, 32, 8, 1, 1, 0, False, False, 0),        # conv_context 1x1 over T words
    (1, 40, 40, 8, 8, 3, 1, 1, True, False, 0),        # up2x, M not a tile multiple
]


def _torch_conv(x, w, b, stride, pad, up2x, act):
    if up2x:
        x = F.interpolate(x, scale_factor=2, mode="nearest")
    y = F.conv2d(x, w, b, stride, pad)
    if act == 2:
        y = F.leaky_relu(y, 0.2)
    elif act == 4:
        y = torch.tanh(y)
    elif act == 5:
        y = torch.sigmoid(y)
    return y


@pytest.mark.parametrize("case", CONV_CASES, ids=lambda c: "x".join(str(int(v)) for v in c))
def test_conv_fwd_bwd(case):
    ops = _ops()
    N, H, W, Ci, Co, k, s, p, up, has_b, act = case
    x = rnd(N, Ci, H, W, seed=1).requires_grad_(True)
    w = rnd(Co, Ci, k, k, seed=2, scale=1.0 / np.sqrt(Ci * k * k)).requires_grad_(True)
    b = rnd(Co, seed=3, scale=0.1).requires_grad_(True) if has_b else None
    y_ref = _torch_conv(x, w, b, s, p, up, act)
    g = rnd(*y_ref.shape, seed=4)
    y_ref.backward(g)

    xd = x.detach().cuda().permute(0, 2, 3, 1).contiguous().requires_grad_(True)
    wd = w.detach().cuda().requires_grad_(True)
    bd = b.detach().cuda().requires_grad_(True) if has_b else None
    y = ops.conv2d(xd, wd, bd, s, p, up, act)
    assert rel(y.permute(0, 3, 1, 2), y_ref) < FP32_TOL
    y.backward(g.cuda().permute(0, 2, 3, 1).contiguous())
    assert rel(xd.grad.permute(0, 3, 1, 2), x.grad) < FP32_TOL
    assert rel(wd.grad, w.grad) < FP32_TOL
    if has_b:
        assert rel(bd.grad, b.grad) < FP32_TOL


@pytest.mark.parametrize("act,name", [(0, "none"), (1, "relu"), (2, "lrelu"), (3, "glu")])
@pytest.mark.parametrize("S", [1, 3])
@pytest.mark.parametrize("res", [False, True])
def test_bn_act(act, name, S, res):
    ops = _ops()
    Bseg, H, W, Cc = 4, 5, 7, 24
    x = rnd(S * Bseg, Cc, H, W, seed=5, scale=2.0) + 0.5
    x.requires_grad_(True)
    bn = torch.nn.BatchNorm2d(Cc)
    with torch.no_grad():
        bn.weight.copy_(1 + 0.1 * rnd(Cc, seed=6))
        bn.bias.copy_(0.1 * rnd(Cc, seed=7))
    Co = Cc // 2 if act == 3 else Cc
    r = rnd(S * Bseg, Co, H, W, seed=8).requires_grad_(True) if res else None
    outs = []
    for s in range(S):  # the reference calls the module once per object
        z = bn(x[s * Bseg:(s + 1) * Bseg])
        z = {0: lambda t: t, 1: F.relu, 2: lambda t: F.leaky_relu(t, 0.2), 3: O.glu}[act](z)
        outs.append(z)
    y_ref = torch.cat(outs, 0)
    if res:
        y_ref = y_ref + r
    g = rnd(*y_ref.shape, seed=9)
    y_ref.backward(g)

    bn2 = torch.nn.BatchNorm2d(Cc).cuda()
    with torch.no_grad():
        bn2.weight.copy_(bn.weight.detach())
        bn2.bias.copy_(bn.bias.detach())
    xd = x.detach().cuda().permute(0, 2, 3, 1).contiguous().requires_grad_(True)
    rd = r.detach().cuda().permute(0, 2, 3, 1).contiguous().requires_grad_(True) if res else None
    y = ops.bn_act(xd, bn2, act, residual=rd, segments=S)
    assert rel(y.permute(0, 3, 1, 2), y_ref) < FP32_TOL
    y.backward(g.cuda().permute(0, 2, 3, 1).contiguous())
    assert rel(xd.grad.permute(0, 3, 1, 2), x.grad) < 1e-4
    assert rel(bn2.weight.grad, bn.weight.grad) < 1e-4
    assert rel(bn2.bias.grad, bn.bias.grad) < 1e-4
    if res:
        assert rel(rd.grad.permute(0, 3, 1, 2), r.grad) < 1e-6
    assert rel(bn2.running_mean, bn.running_mean) < 1e-5
    assert rel(bn2.running_var, bn.running_var) < 1e-5
    assert int(bn2.num_batches_tracked) == S


def test_bn1d_many_channels():
    """BatchNorm1d over 24576 features of 32 rows (INIT_STAGE_G.fc) + GLU."""
    ops = _ops()
    x = rnd(32, 4096, seed=1).requires_grad_(True)
    bn = torch.nn.BatchNorm1d(4096)
    y_ref = O.glu(bn(x))
    g = rnd(*y_ref.shape, seed=2)
    y_ref.backward(g)
    bn2 = torch.nn.BatchNorm1d(4096).cuda()
    xd = x.detach().cuda().requires_grad_(True)
    y = ops.bn_act(xd, bn2, 3)
    assert rel(y, y_ref) < FP32_TOL
    y.backward(g.cuda())
    assert rel(xd.grad, x.grad) < 1e-4


def test_plain_glu():
    ops = _ops()
    x = rnd(6, 40, seed=3).requires_grad_(True)
    y_ref = O.glu(x)
    g = rnd(6, 20, seed=4)
    y_ref.backward(g)
    xd = x.detach().cuda().requires_grad_(True)
    y = ops.activation(xd, 3)
    y.backward(g.cuda())
    assert rel(y, y_ref) < 1e-6 and rel(xd.grad, x.grad) < 1e-6


@pytest.mark.parametrize("tag", ["scatter16", "crop64to16", "scatter15to16", "crop256to32"])
def test_stn_golden(tag):
    """Single-object STN against vectors from the reference's own stn() (model.py:17-21)."""
    ops = _ops()
    G, _ = gu.load("stn_cases")
    theta = gu.full(G, tag + "/theta").cuda()
    shp = G[tag + "/y"]["shape"]
    if "full" in G[tag + "/x"]:
        x = gu.full(G, tag + "/x")
    else:  # large input not stored: regenerate exactly like make_golden.stn_cases
        rng = np.random.RandomState(11)
        for t2, ish in (("scatter16", (4, 6, 16, 16)), ("crop64to16", (4, 3, 64, 64)),
                        ("scatter15to16", (4, 5, 15, 15)), ("crop256to32", (4, 3, 256, 256))):
            oshape = G[t2 + "/y"]["shape"]
            xx = rng.standard_normal(ish).astype(np.float32)
            rng.standard_normal(oshape)
            if t2 == tag:
                x = torch.from_numpy(xx)
                break
        gu.check(x, G[tag + "/x"], 1e-7, "regenerated x")
    B = x.shape[0]
    xd = x.cuda().permute(0, 2, 3, 1).contiguous().requires_grad_(True)
    y = ops.stn_scatter_sum(xd, theta.reshape(B, 1, 2, 3), B, 1, (shp[2], shp[3]))
    gu.check(y.permute(0, 3, 1, 2), G[tag + "/y"], 1e-5, tag)
    y.backward(gu.full(G, tag + "/g").cuda().permute(0, 2, 3, 1).contiguous())
    gu.check(xd.grad.permute(0, 3, 1, 2), G[tag + "/dx"], 1e-5, tag + " dx")
    if tag == "scatter16":
        assert float(y[1].abs().max()) == 0.0  # empty slot -> exact zeros


def test_stn_fused_objects():
    """scatter-sum over 3 objects and crop+label-concat against the oracle's per-object loops."""
    ops = _ops()
    from mog_b200 import synth
    rng = np.random.RandomState(3)
    B, S, Cc = 5, 3, 12
    _, _, onehot, theta, theta_inv = synth.bboxes_and_labels(rng, B)
    theta, theta_inv, onehot = map(torch.from_numpy, (theta, theta_inv, onehot))
    x = rnd(S, B, Cc, 15, 15, seed=1).requires_grad_(True)
    y_ref = sum(O.stn(x[s], theta_inv[:, s], (B, Cc, 16, 16)) for s in range(S))
    g = rnd(*y_ref.shape, seed=2)
    y_ref.backward(g)
    xd = x.detach().reshape(S * B, Cc, 15, 15).cuda().permute(0, 2, 3, 1).contiguous().requires_grad_(True)
    y = ops.stn_scatter_sum(xd, theta_inv.cuda(), B, S, (16, 16))
    assert rel(y.permute(0, 3, 1, 2), y_ref) < 1e-5
    y.backward(g.cuda().permute(0, 2, 3, 1).contiguous())
    assert rel(xd.grad.permute(0, 3, 1, 2).reshape(S, B, Cc, 15, 15), x.grad) < 1e-5
    # crop + concat
    img = rnd(B, 3, 64, 64, seed=5).requires_grad_(True)
    refs = []
    for s in range(S):
        patch = O.stn(img, theta[:, s], (B, 3, 16, 16))
        lab = onehot[:, s].reshape(B, 81, 1, 1).repeat(1, 1, 16, 16)
        refs.append(torch.cat((patch, lab), 1))
    c_ref = torch.cat(refs, 0)
    g2 = rnd(*c_ref.shape, seed=6)
    c_ref.backward(g2)
    imd = img.detach().cuda().permute(0, 2, 3, 1).contiguous().requires_grad_(True)
    c = ops.stn_crop(imd, theta.cuda(), S, (16, 16), extra=onehot.cuda())
    assert rel(c.permute(0, 3, 1, 2), c_ref) < 1e-5
    c.backward(g2.cuda().permute(0, 2, 3, 1).contiguous())
    assert rel(imd.grad.permute(0, 3, 1, 2), img.grad) < 1e-5


@pytest.mark.parametrize("tag", ["b3", "b4"])
def test_word_attention_golden(tag):
    """Fused attention (+1x1 conv) against the reference's GlobalAttentionGeneral, incl. the
    mask tiling quirk with B not dividing queryL (b3)."""
    from mog_b200.attngan.GlobalAttention import GlobalAttentionGeneral
    G, _ = gu.load("attention_cases")
    h, ctx, w = gu.full(G, tag + "/h"), gu.full(G, tag + "/ctx"), gu.full(G, tag + "/w")
    mask = gu.full(G, tag + "/mask").bool()
    att = GlobalAttentionGeneral(w.shape[0], w.shape[1]).cuda()
    with torch.no_grad():
        att.conv_context.weight.copy_(w)
    hd = h.cuda().requires_grad_(True)
    cd = ctx.cuda().requires_grad_(True)
    att.applyMask(mask.cuda())
    wc, attn = att(hd, cd)
    gu.check(wc, G[tag + "/wc"], 1e-5, "wc")
    gu.check(attn, G[tag + "/attn"], 1e-5, "attn")
    wc.backward(gu.full(G, tag + "/g").cuda())
    gu.check(hd.grad, G[tag + "/dh"], 2e-5, "dh")
    gu.check(cd.grad, G[tag + "/dctx"], 2e-5, "dctx")
    gu.check(att.conv_context.weight.grad, G[tag + "/dw"], 2e-5, "dw")


def test_sigmoid_bce():
    ops = _ops()
    z = (rnd(37, seed=1) * 4).requires_grad_(True)
    z.data[0], z.data[1] = 200.0, -200.0  # saturated: exercises the log clamp at -100
    for tval in (1.0, 0.0):
        t = torch.full((37,), tval)
        ref = F.binary_cross_entropy(torch.sigmoid(z), t)
        (gz,) = torch.autograd.grad(ref * 0.7, z)
        zd = z.detach().cuda().requires_grad_(True)
        out = ops.sigmoid_bce(zd, t.cuda())
        (out * 0.7).backward()
        assert rel(out, ref) < 1e-6
        assert rel(zd.grad, gz) < 1e-5


def test_abi_errors_are_loud():
    """Bad arguments return negative codes with a message; CPU tensors are refused (no fallback)."""
    import ctypes as C
    from mog_b200 import _lib, ops
    L = _lib.lib()
    assert L.mog_version() >= 100
    rc = L.mog_stn_fwd(None, None, None, None, 0, 1, 1, 1, 1, 1, 1, 1, 1, 0, None)
    assert rc == -1 and b"null" in L.mog_last_error()
    d = _lib.MogConvDesc(1, 4, 4, 8, 8, 9, 9, 1, 0, 0, 0, 0)
    assert L.mog_conv_out_hw(C.byref(d), None, None) == -1
    with pytest.raises(RuntimeError):
        ops.conv2d(torch.zeros(1, 4, 4, 8), torch.zeros(8, 8, 3, 3))


def test_damsm_losses_golden():
    """Fused words_loss (+ sent_loss) against the reference's miscc/losses.py on synthetic features."""
    from mog_b200.attngan.miscc import losses as L
    from mog_b200.attngan.miscc.config import cfg, reset_cfg
    reset_cfg()
    cfg.TRAIN.SMOOTH.GAMMA1, cfg.TRAIN.SMOOTH.GAMMA2, cfg.TRAIN.SMOOTH.GAMMA3 = 4.0, 5.0, 10.0
    G, _ = gu.load("attention_cases")
    feat = gu.full(G, "damsm/feat").cuda().requires_grad_(True)
    code = gu.full(G, "damsm/code").cuda().requires_grad_(True)
    words, sent = gu.full(G, "damsm/words").cuda(), gu.full(G, "damsm/sent").cuda()
    lens = gu.full(G, "damsm/lens").long()
    B = feat.shape[0]
    labels = torch.arange(B, device="cuda")
    w0, w1, _ = L.words_loss(feat, words, labels, lens, np.arange(B), B)
    s0, s1 = L.sent_loss(code, sent, labels, np.arange(B), B)
    for k, v in {"w0": w0, "w1": w1, "s0": s0, "s1": s1}.items():
        gu.check(v, G["damsm/" + k], 2e-5, k)
    (w0 + w1 + s0 + s1).backward()
    gu.check(feat.grad, G["damsm/dfeat"], 1e-4, "dfeat")
    gu.check(code.grad, G["damsm/dcode"], 1e-4, "dcode")
    reset_cfg()


def test_func_attention_matches_oracle():
    from mog_b200.attngan.GlobalAttention import func_attention
    q = rnd(3, 32, 7, seed=1)
    c = rnd(3, 32, 17, 17, seed=2)
    wref, aref = O.func_attention(q, c, 4.0)
    w, a = func_attention(q.cuda(), c.cuda(), 4.0)
    assert rel(w, wref) < 1e-5 and rel(a, aref) < 1e-5
