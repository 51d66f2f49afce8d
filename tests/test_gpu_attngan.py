"""Network-level parity of the libmog AttnGAN modules (G_NET, D_NET64/128/256, losses) against
(a) golden vectors produced by executing the unmodified reference and (b) the CPU oracle on a
second seed.  Run on the B200 box: -m gpu."""
import json
import os

import numpy as np
import pytest
import torch

import golden_util as gu
from mog_b200 import synth
from oracle import attngan_oracle as O

pytestmark = pytest.mark.gpu

# fp32 CUDA-core path vs fp32 CPU reference: summation order only.  Gradients of deep nets
# accumulate a little more.
OUT_TOL = 5e-5
GRAD_TOL = 5e-4


def _set_cfg(c):
    from mog_b200.attngan.miscc.config import cfg, reset_cfg
    reset_cfg()
    cfg.GAN.GF_DIM, cfg.GAN.DF_DIM, cfg.GAN.Z_DIM = c["GF_DIM"], c["DF_DIM"], c["Z_DIM"]
    cfg.GAN.R_NUM, cfg.TEXT.EMBEDDING_DIM, cfg.TEXT.WORDS_NUM = c["R_NUM"], c["EMBEDDING_DIM"], c["T"]
    cfg.TRAIN.BATCH_SIZE = c["B"]
    return cfg


def _build(c, seed):
    from mog_b200.attngan import model as M
    _set_cfg(c)
    netG = M.G_NET()
    netsD = [M.D_NET64(), M.D_NET128(), M.D_NET256()]
    netG.load_state_dict(synth.fill_state_dict(netG.state_dict(), seed + 1))
    for i, d in enumerate(netsD):
        d.load_state_dict(synth.fill_state_dict(d.state_dict(), seed + 2 + i))
    return netG.cuda().train(), [d.cuda().train() for d in netsD]


def _dev(batch):
    out = {}
    for k, v in batch.items():
        if isinstance(v, list):
            out[k] = [t.cuda() for t in v]
        elif torch.is_tensor(v):
            out[k] = v.cuda()
        else:
            out[k] = v
    return out


def test_attngan_step_vs_reference_golden():
    """G forward, 3 discriminator losses + all D gradients + BN running stats against the
    reference run (tests/golden/attngan_tiny_step), then G adversarial+KL gradients
    (tests/golden/attngan_tiny_gd)."""
    from mog_b200.attngan.miscc import losses as L
    G, meta = gu.load("attngan_tiny_step")
    G2, _ = gu.load("attngan_tiny_gd")
    c, seed = meta["cfg"], meta["seed"]
    netG, netsD = _build(c, seed)
    b = _dev(synth.attngan_batch(c["B"], T=c["T"], nef=c["EMBEDDING_DIM"], nz=c["Z_DIM"], seed=seed))
    eps = gu.full(G, "G/eps").cuda()
    B = c["B"]
    real, fake = torch.ones(B, device="cuda"), torch.zeros(B, device="cuda")
    tm, tmi, oh = b["transf_matrices"], b["transf_matrices_inv"], b["label_one_hot"]
    imgs, atts, mu, logvar = netG(b["noise"], b["sent_emb"], b["words_embs"], b["mask"], tmi, oh, eps=eps)
    for i in range(3):
        gu.check(imgs[i], G["G/fake%d" % i], OUT_TOL, "fake%d" % i)
    for i in range(2):
        gu.check(atts[i], G["G/att%d" % i], OUT_TOL, "att%d" % i)
    gu.check(mu, G["G/mu"], OUT_TOL, "mu")
    gu.check(logvar, G["G/logvar"], OUT_TOL, "logvar")
    for i, netD in enumerate(netsD):
        netD.zero_grad()
        kw = dict(local_labels=oh, transf_matrices=tm, transf_matrices_inv=tmi) if i == 0 else {}
        errD = L.discriminator_loss(netD, b["imgs"][i], imgs[i], b["sent_emb"], real, fake, [0], **kw)
        errD.backward()
        gu.check(errD, G["D%d/errD" % i], OUT_TOL, "errD%d" % i)
        for k, p in netD.named_parameters():
            gu.check(p.grad, G["D%d/grad/%s" % (i, k)], GRAD_TOL, "D%d grad %s" % (i, k))
        for k, v in netD.state_dict().items():
            if "running" in k:
                gu.check(v, G["D%d/buf_after_dstep/%s" % (i, k)], OUT_TOL, k)
    # generator step, adversarial + KL (no DAMSM), D weights frozen (their wgrad is skipped)
    for d in netsD:
        for p in d.parameters():
            p.requires_grad_(False)
    netG.zero_grad()
    errG, _ = L.generator_loss(netsD, None, imgs, real, b["words_embs"], b["sent_emb"], None, b["cap_lens"],
                               b["class_ids"], [0], local_labels=oh, transf_matrices=tm, transf_matrices_inv=tmi)
    kl = L.KL_loss(mu, logvar)
    gu.check(errG, G2["G/errG_adv"], OUT_TOL, "errG")
    gu.check(kl, G2["G/kl"], OUT_TOL, "kl")
    (errG + kl).backward()
    for k, p in netG.named_parameters():
        gu.check(p.grad, G2["G/grad/%s" % k], GRAD_TOL, "G grad %s" % k)


def test_attngan_step_vs_oracle_second_seed():
    """Same step on different data/weights (odd batch, different T) against the CPU oracle."""
    from mog_b200.attngan.miscc import losses as L
    c = dict(GF_DIM=8, DF_DIM=8, Z_DIM=16, R_NUM=1, EMBEDDING_DIM=24, T=9, B=3)
    seed = 321
    netG, netsD = _build(c, seed)
    cfgo = O.Cfg(GF_DIM=8, DF_DIM=8, Z_DIM=16, R_NUM=1, EMBEDDING_DIM=24)
    PG = O.leafify(netG.state_dict())
    PDs = [O.leafify(d.state_dict()) for d in netsD]
    batch = synth.attngan_batch(c["B"], T=c["T"], nef=c["EMBEDDING_DIM"], nz=c["Z_DIM"], seed=seed)
    ref = O.gd_step(PG, PDs, cfgo, batch)
    b = _dev(batch)
    B = c["B"]
    real, fake = torch.ones(B, device="cuda"), torch.zeros(B, device="cuda")
    tm, tmi, oh = b["transf_matrices"], b["transf_matrices_inv"], b["label_one_hot"]
    imgs, _, mu, logvar = netG(b["noise"], b["sent_emb"], b["words_embs"], b["mask"], tmi, oh, eps=b["eps"])
    for i in range(3):
        assert gu.rel_l2(imgs[i].detach().cpu().numpy(), ref["fake_imgs"][i].numpy()) < OUT_TOL
    for i, netD in enumerate(netsD):
        kw = dict(local_labels=oh, transf_matrices=tm, transf_matrices_inv=tmi) if i == 0 else {}
        errD = L.discriminator_loss(netD, b["imgs"][i], imgs[i], b["sent_emb"], real, fake, [0], **kw)
        errD.backward()
        assert abs(float(errD) - float(ref["errD"][i])) < 1e-4 * abs(float(ref["errD"][i]))
        for k, p in netD.named_parameters():
            assert gu.rel_l2(p.grad.cpu().numpy(), PDs[i][k].grad.numpy()) < GRAD_TOL, k
    for d in netsD:
        for p in d.parameters():
            p.requires_grad_(False)
    errG, _ = L.generator_loss(netsD, None, imgs, real, b["words_embs"], b["sent_emb"], None, b["cap_lens"],
                               b["class_ids"], [0], local_labels=oh, transf_matrices=tm, transf_matrices_inv=tmi)
    (errG + L.KL_loss(mu, logvar)).backward()
    assert abs(float(errG) - float(ref["errG"])) < 1e-4 * abs(float(ref["errG"]))
    for k, p in netG.named_parameters():
        assert gu.rel_l2(p.grad.cpu().numpy(), PG[k].grad.numpy()) < GRAD_TOL, k


def test_public_api_shapes_and_reference_signatures():
    """Reference-facing surface: NCHW in/out, COND/UNCOND heads return probabilities."""
    c = dict(GF_DIM=8, DF_DIM=8, Z_DIM=16, R_NUM=1, EMBEDDING_DIM=24, T=9, B=2)
    netG, netsD = _build(c, 5)
    b = _dev(synth.attngan_batch(2, T=9, nef=24, nz=16, seed=5))
    imgs, atts, mu, logvar = netG(b["noise"], b["sent_emb"], b["words_embs"], b["mask"],
                                  b["transf_matrices_inv"], b["label_one_hot"])
    assert [tuple(i.shape) for i in imgs] == [(2, 3, 64, 64), (2, 3, 128, 128), (2, 3, 256, 256)]
    assert [tuple(a.shape) for a in atts] == [(2, 9, 64, 64), (2, 9, 128, 128)]
    assert mu.shape == (2, 100) and logvar.shape == (2, 100)
    f = netsD[0](b["imgs"][0], b["label_one_hot"], b["transf_matrices"], b["transf_matrices_inv"])
    assert tuple(f.shape) == (2, 64, 4, 4)
    # NCHW-contiguous input (as a torch DataLoader would deliver) is converted, not rejected
    f2 = netsD[2](b["imgs"][2].contiguous())
    assert tuple(f2.shape) == (2, 64, 4, 4)
    p = netsD[0].COND_DNET(f, b["sent_emb"])
    q = netsD[0].UNCOND_DNET(f)
    assert p.shape == (2,) and q.shape == (2,) and float(p.min()) > 0 and float(p.max()) < 1


@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
def test_train_step_matches_reference(precision):
    """a20 -- the function bench.py times: three consecutive ``condGANTrainer.train_step`` calls (G forward, per D:
    pair pass + backward + fused Adam + repack, G loss incl. the DAMSM branch, backward, fused Adam + EMA) against the
    reference's trainer.py:294-342 executed with ``optim.Adam`` (tests/golden/make_golden_trainstep.py): parameters,
    EMA copy, Adam moments and BatchNorm buffers after the third step, the losses of every step."""
    from mog_b200 import ops
    from mog_b200.attngan import model as M
    from mog_b200.attngan.trainer import condGANTrainer
    from test_oracle_golden import check_train_state
    G, meta = gu.load("attngan_tiny_trainstep")
    c, seed, K = meta["cfg"], meta["seed"], meta["steps"]
    cfg = _set_cfg(c)
    cfg.TRAIN.SMOOTH.GAMMA1, cfg.TRAIN.SMOOTH.GAMMA2, cfg.TRAIN.SMOOTH.GAMMA3, cfg.TRAIN.SMOOTH.LAMBDA = 4.0, 5.0, 10.0, 50.0
    old = ops.get_precision()
    cfg.MOG.PRECISION = precision          # (the trainer's constructor applies cfg.MOG.PRECISION)
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False      # the stand-in encoder is a torch conv (test infrastructure)
    try:
        netG = M.G_NET()
        netsD = [M.D_NET64(), M.D_NET128(), M.D_NET256()]
        netG.load_state_dict(synth.fill_state_dict(netG.state_dict(), seed + 1))
        for i, d in enumerate(netsD):
            d.load_state_dict(synth.soften_logits(synth.fill_state_dict(d.state_dict(), seed + 2 + i), meta["logit_scale"]))
        netG.cuda().train()
        for d in netsD:
            d.cuda().train()
        tr = condGANTrainer("", None, 0, None)
        assert ops.get_precision() == ops.PREC_NAMES[precision]
        tr.image_encoder = synth.StandInEncoder(c["EMBEDDING_DIM"], device="cuda")
        optG, optDs = tr.define_optimizers(netG, netsD)
        st = tr.make_step_state(netG, netsD, optG, optDs)
        b = _dev(synth.attngan_batch(c["B"], T=c["T"], nef=c["EMBEDDING_DIM"], nz=c["Z_DIM"], seed=seed))
        tol = 2e-4 if precision == "fp32" else 5e-4
        for k in range(K):
            noise = torch.from_numpy(np.random.RandomState(seed + 10 + k).standard_normal((c["B"], c["Z_DIM"])).astype(np.float32)).cuda()
            errD, errG, kl = tr.train_step(st, b["imgs"], b["sent_emb"], b["words_embs"], b["mask"], b["transf_matrices"],
                                           b["transf_matrices_inv"], b["label_one_hot"], b["cap_lens"], b["class_ids"],
                                           noise=noise, eps=gu.full(G, "step%d/eps" % k).cuda())
            gu.check(errD, G["step%d/errD_total" % k], tol, "errD step %d" % k)
            gu.check(errG, G["step%d/errG_total" % k], tol, "errG step %d" % k)
            gu.check(kl, G["step%d/kl" % k], tol, "kl step %d" % k)
        nets = {"G": (dict(netG.state_dict(keep_vars=True)), optG.state)}
        for i, d in enumerate(netsD):
            nets["D%d" % i] = (dict(d.state_dict(keep_vars=True)), optDs[i].state)
        ema = dict(zip([n for n, _ in netG.named_parameters()], st["avg_param_G"]))
        check_train_state(G, nets, ema, scale=1.0 if precision == "fp32" else 2.0)
    finally:
        ops.set_precision(old)
        torch.backends.cudnn.allow_tf32 = tf32


def test_graphed_step_equals_eager_steps():
    """The CUDA-graph form of the step (``condGANTrainer.graphed_step``: what bench.py and ``train()`` replay) against the
    eager ``train_step`` on the same noise / eps / batches: parameters, EMA and Adam state after 2 warm-up + 3 replayed
    steps (device-resident Adam step counter, repacked weights, BatchNorm running statistics all advance inside the graph)."""
    from mog_b200 import ops
    from mog_b200.attngan import model as M
    from mog_b200.attngan.trainer import condGANTrainer
    c = dict(GF_DIM=8, DF_DIM=8, Z_DIM=20, R_NUM=1, EMBEDDING_DIM=32, T=6, B=4)
    cfg = _set_cfg(c)
    cfg.MOG.PRECISION = "bf16x3"
    seed, K = 77, 5
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False

    def make():
        netG, netsD = _build(c, seed)
        tr = condGANTrainer("", None, 0, None)
        tr.image_encoder = synth.StandInEncoder(c["EMBEDDING_DIM"], device="cuda")
        optG, optDs = tr.define_optimizers(netG, netsD)
        return tr, tr.make_step_state(netG, netsD, optG, optDs)

    try:
        b = _dev(synth.attngan_batch(c["B"], T=c["T"], nef=c["EMBEDDING_DIM"], nz=c["Z_DIM"], seed=seed))
        rng = np.random.RandomState(5)
        noises = [torch.from_numpy(rng.standard_normal((c["B"], c["Z_DIM"])).astype(np.float32)).cuda() for _ in range(K)]
        epss = [torch.from_numpy(rng.standard_normal((c["B"], 100)).astype(np.float32)).cuda() for _ in range(K)]
        args = (b["imgs"], b["sent_emb"], b["words_embs"], b["mask"], b["transf_matrices"], b["transf_matrices_inv"], b["label_one_hot"],
                b["cap_lens"], b["class_ids"])
        tr1, st1 = make()
        losses1 = [tr1.train_step(st1, *args, noise=noises[k], eps=epss[k]) for k in range(K)]
        tr2, st2 = make()
        # two eager steps, then three replays of the captured graph
        gs = None
        losses2 = []
        for k in range(K):
            if k < 2:
                losses2.append(tr2.train_step(st2, *args, noise=noises[k], eps=epss[k]))
            else:
                if gs is None:     # capture without training (dry warm-up): the replays continue where the eager steps stopped
                    gs = tr2.graphed_step(st2, *args, warmup=1, noise=noises[k], eps=epss[k], dry_warmup=True)
                losses2.append([t.clone() for t in gs(*args, noise=noises[k], eps=epss[k])])
        torch.cuda.synchronize()
        # (gates as in the reference comparison; at the benchmark's size the same comparison is made bit for bit:
        # tests/test_gpu_fullsize.py)
        rel = lambda a, r: float((a.double() - r.double()).norm() / r.double().norm().clamp_min(1e-30))   # noqa: E731
        for k in range(K):
            for a, r in zip(losses2[k], losses1[k]):
                assert abs(float(a) - float(r)) <= 1e-4 * abs(float(r)) + 1e-6, (k, float(a), float(r))
        moved = 0.0
        for n1, n2 in zip([st1["netG"]] + st1["netsD"], [st2["netG"]] + st2["netsD"]):
            num = den = 0.0
            for (name, p1), (_, p2) in zip(n1.named_parameters(), n2.named_parameters()):
                assert rel(p2, p1) <= 1e-3, (name, rel(p2, p1))
                num += float((p1.double() - p2.double()).pow(2).sum()); den += float(p1.double().pow(2).sum())
            assert (num / den) ** 0.5 <= 2e-4
            for (name, v1), (_, v2) in zip(n1.named_buffers(), n2.named_buffers()):
                assert rel(v2.float(), v1.float()) <= 2e-3, name
        for a1, a2 in zip(st1["avg_param_G"], st2["avg_param_G"]):
            assert rel(a2, a1) <= 1e-5
        # the replays really stepped: Adam counts advanced to K on host and device, the weights moved since capture
        for o1, o2 in zip([st1["optG"]] + st1["optDs"], [st2["optG"]] + st2["optDs"]):
            s1, s2 = o1.state_dict()["state"], o2.state_dict()["state"]
            for i in s1:
                assert float(s1[i]["step"]) == float(s2[i]["step"]) == K
            for slot in o2._dev_steps.values():
                assert float(slot[0]) == K and slot[1] == K
    finally:
        torch.backends.cudnn.allow_tf32 = tf32


def test_graphed_step_without_damsm_captures():
    """The G+D-only step (no image encoder: ``bench.py --no-damsm``) as a CUDA graph with the branch streams on: every forked
    stream is joined and no un-forked one is waited on (a capture error otherwise)."""
    from mog_b200.attngan.trainer import condGANTrainer
    c = dict(GF_DIM=8, DF_DIM=8, Z_DIM=20, R_NUM=1, EMBEDDING_DIM=32, T=6, B=4)
    cfg = _set_cfg(c)
    cfg.MOG.PRECISION = "bf16x3"
    netG, netsD = _build(c, 31)
    tr = condGANTrainer("", None, 0, None)
    tr.image_encoder = None
    optG, optDs = tr.define_optimizers(netG, netsD)
    st = tr.make_step_state(netG, netsD, optG, optDs)
    b = _dev(synth.attngan_batch(c["B"], T=c["T"], nef=c["EMBEDDING_DIM"], nz=c["Z_DIM"], seed=31))
    args = (b["imgs"], b["sent_emb"], b["words_embs"], b["mask"], b["transf_matrices"], b["transf_matrices_inv"], b["label_one_hot"],
            b["cap_lens"], b["class_ids"])
    gs = tr.graphed_step(st, *args, warmup=2)
    out = [float(t) for t in gs(*args)]
    torch.cuda.synchronize()
    assert all(np.isfinite(v) for v in out)
