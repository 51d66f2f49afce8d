"""Fused Adam(+EMA) (mog_adam_multi) against torch.optim.Adam and the reference's EMA loop."""
import pytest
import torch


def _params(seed, device):
    g = torch.Generator().manual_seed(seed)
    shapes = [(96, 48, 3, 3), (7,), (1,), (33, 5), (257,), (4096 * 5 + 3,), (768, 100)]
    return [torch.randn(s, generator=g).to(device).requires_grad_(True) for s in shapes]


@pytest.mark.gpu
@pytest.mark.parametrize("with_ema", [False, True])
def test_adam_matches_torch(with_ema):
    from mog_b200.optim import Adam
    pa, pb = _params(1, "cuda"), _params(1, "cuda")
    oa = torch.optim.Adam(pa, lr=2e-4, betas=(0.5, 0.999))
    ob = Adam(pb, lr=2e-4, betas=(0.5, 0.999))
    ea = [p.detach().clone() for p in pa]
    eb = [p.detach().clone() for p in pb]
    g = torch.Generator().manual_seed(2)
    for it in range(5):
        for a, b in zip(pa, pb):
            gr = torch.randn(a.shape, generator=g).cuda() * (10.0 ** (it - 3))
            a.grad, b.grad = gr.clone(), gr.clone()
        oa.step()
        if with_ema:
            for avg, p in zip(ea, pa):   # attngan/trainer.py:341-342
                avg.mul_(0.999).add_(p.data, alpha=0.001)
            ob.step(ema_params=eb)
        else:
            ob.step()
    for a, b in zip(pa, pb):
        assert torch.allclose(a, b, rtol=2e-6, atol=1e-7), (a - b).abs().max()
    for a, b in zip(ea, eb):
        assert torch.allclose(a, b, rtol=2e-6, atol=1e-7)
    sa, sb = oa.state_dict(), ob.state_dict()
    for k in sa["state"]:
        assert float(sa["state"][k]["step"]) == float(sb["state"][k]["step"]) == 5
        assert torch.allclose(sa["state"][k]["exp_avg"], sb["state"][k]["exp_avg"], rtol=2e-6, atol=1e-9)
        assert torch.allclose(sa["state"][k]["exp_avg_sq"], sb["state"][k]["exp_avg_sq"], rtol=2e-6, atol=1e-12)
    # state_dicts interchange: continue the torch run from the libmog state
    oc = torch.optim.Adam(_params(1, "cuda"), lr=2e-4, betas=(0.5, 0.999))
    oc.load_state_dict(sb)
    assert float(oc.state_dict()["state"][0]["step"]) == 5


@pytest.mark.gpu
def test_adam_unaligned_views_and_grad_scale():
    """Parameters that are views into a flat bucket at odd offsets take the scalar path; grad_scale = 1/world."""
    from mog_b200.optim import Adam
    flat = torch.randn(1000, device="cuda")
    ref = flat.clone()
    views = [flat[1:8].detach().requires_grad_(True), flat[9:510].detach().requires_grad_(True)]
    rviews = [ref[1:8].detach().clone().requires_grad_(True), ref[9:510].detach().clone().requires_grad_(True)]
    ob, oa = Adam(views, lr=1e-2, betas=(0.5, 0.999)), torch.optim.Adam(rviews, lr=1e-2, betas=(0.5, 0.999))
    for v, r in zip(views, rviews):
        gr = torch.randn_like(r)
        v.grad, r.grad = (gr * 4).clone(), gr.clone()
    ob.step(grad_scale=0.25)
    oa.step()
    for v, r in zip(views, rviews):
        assert torch.allclose(v, r, rtol=2e-6, atol=1e-7)
    assert torch.equal(flat[0], ref[0]) and torch.equal(flat[8], ref[8]) and torch.equal(flat[510:], ref[510:])


def test_adam_refuses_cpu_params():
    from mog_b200.optim import Adam
    p = torch.zeros(4, requires_grad=True)
    p.grad = torch.ones(4)
    with pytest.raises(RuntimeError):
        Adam([p], lr=1e-3).step()


@pytest.mark.gpu
@pytest.mark.parametrize("prec", ["fp32", "bf16x3"])
def test_adam_update_reaches_the_next_forward(prec):
    """The fused kernel writes parameters through raw pointers (torch's version counter does not move): the packed-weight
    cache of the convolutions must still notice, i.e. the forward after step() uses the NEW weights."""
    import torch.nn.functional as F
    from mog_b200 import ops
    from mog_b200.optim import Adam
    torch.backends.cudnn.allow_tf32 = False
    torch.manual_seed(11)
    w = (torch.randn(32, 16, 3, 3, device="cuda") * 0.1).requires_grad_(True)
    x = torch.randn(2, 12, 12, 16, device="cuda")
    opt = Adam([w], lr=0.05, betas=(0.5, 0.999))
    P = ops.PREC_NAMES[prec]
    for it in range(3):
        y = ops.conv2d(x, w, None, 1, 1, False, ops.ACT_NONE, P)
        ref = F.conv2d(x.permute(0, 3, 1, 2), w.detach(), None, 1, 1).permute(0, 2, 3, 1)
        err = float((y.detach() - ref).norm() / ref.norm())
        assert err < 1e-4, (it, err)          # a stale pack would be off by O(lr) = percent
        opt.zero_grad(set_to_none=True)
        y.square().mean().backward()
        w_before = w.detach().clone()
        opt.step()
        assert float((w.detach() - w_before).abs().max()) > 1e-3


REPACK_CASES = [
    # Cout, Cin, k, stride, pad, up2x, H      (x is [2, H, H, Cin])
    (192, 96, 3, 1, 1, False, 16),    # ResBlock conv: fwd + dgrad (transposed) packs
    (96, 96, 3, 1, 1, True, 16),      # upBlock: 4 sub-pixel phase packs with pre-summed 2x2 taps, dgrad through parity views
    (192, 96, 4, 2, 1, False, 16),    # downBlock 4x4/s2: parity-view forward, 4 stride-phase dgrad packs
    (40, 88, 4, 1, 1, False, 16),     # ragged channel counts (Cout 40 -> N tile 48, Cin 88)
    (64, 48, 1, 1, 0, False, 12),     # 1x1
    (1536, 768, 4, 2, 1, False, 8),   # deep discriminator layer: several N tiles, K = 12288
]


@pytest.mark.gpu
@pytest.mark.parametrize("case", REPACK_CASES, ids=lambda c: "x".join(str(int(v)) for v in c))
def test_multi_tensor_repack_equals_per_problem_pack(case):
    """ops.repack (ONE mog_pack_multi launch over all cached layouts of the updated weights -- what runs after every fused Adam
    step) must leave exactly the bytes mog_pack_weight (the lazy per-problem pack) writes."""
    from mog_b200 import ops
    Co, Ci, k, s, p, up, H = case
    P = ops.PREC_NAMES["bf16x3"]
    torch.manual_seed(5)
    w = (torch.randn(Co, Ci, k, k, device="cuda") * 0.05).requires_grad_(True)
    x = torch.randn(2, H, H, Ci, device="cuda", requires_grad=True)

    def fwd_bwd():
        y = ops.conv2d(x, w, None, s, p, up, ops.ACT_NONE, P)
        y.sum().backward()
        x.grad = None
        w.grad = None

    fwd_bwd()                                    # creates the packed layouts (lazy mog_pack_weight)
    cache = w._mog_pack
    assert len(cache) >= 1
    with torch.no_grad():
        w.mul_(1.7).add_(0.01)                   # new values
    ops.invalidate_packed([w])
    for ent in cache.values():                   # (alignment gaps between phase blocks are written by neither kernel)
        ent["out"].zero_()
    ops.repack([w])                              # multi-tensor kernel
    torch.cuda.synchronize()
    multi = {key: ent["out"].clone() for key, ent in cache.items()}
    covered = [key for key, ent in cache.items() if ent["ver"] == ops._weight_version(w)]
    assert covered, "no layout of this weight went through mog_pack_multi"
    for ent in cache.values():                   # force the lazy path to rewrite every layout
        ent["ver"] = None
        ent["out"].zero_()
    fwd_bwd()
    torch.cuda.synchronize()
    for key in covered:
        assert torch.equal(multi[key].view(torch.int32), cache[key]["out"].view(torch.int32)), key
