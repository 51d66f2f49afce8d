"""Pin the oracle (oracle/attngan_oracle.py) against vectors produced by executing the
unmodified reference (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

import golden_util as gu
from mog_b200 import synth
from oracle import attngan_oracle as O

TOL = 2e-5  # fp32 CPU vs fp32 CPU; differences are only op-ordering (e.g. x+0 canvas adds)


def _tiny_cfg(c):
    return O.Cfg(GF_DIM=c["GF_DIM"], DF_DIM=c["DF_DIM"], Z_DIM=c["Z_DIM"], R_NUM=c["R_NUM"],
                 EMBEDDING_DIM=c["EMBEDDING_DIM"])


def _build(keys, seed):
    shapes = {k: torch.empty(s) for k, s in keys.items()}
    return O.leafify(synth.fill_state_dict(shapes, seed))


@pytest.fixture(scope="module")
def keys():
    import json, os
    with open(os.path.join(os.path.dirname(__file__), "golden", "attngan_state_dict_keys.json")) as f:
        return json.load(f)


def test_theta_matches_reference():
    G, _ = gu.load("stn_cases")
    bbox = gu.full(G, "bbox").numpy()
    gu.check(torch.from_numpy(synth.transformation_matrix(bbox)), G["theta_ref"], 1e-7, "theta")
    gu.check(torch.from_numpy(synth.transformation_matrix_inverse(bbox)), G["theta_inv_ref"], 1e-7, "theta_inv")


@pytest.mark.parametrize("tag", ["scatter16", "crop64to16", "scatter15to16"])
def test_stn(tag):
    G, _ = gu.load("stn_cases")
    x = gu.full(G, tag + "/x").requires_grad_(True)
    y = O.stn(x, gu.full(G, tag + "/theta"), G[tag + "/y"]["shape"])
    gu.check(y, G[tag + "/y"], 1e-6, tag)
    y.backward(gu.full(G, tag + "/g"))
    gu.check(x.grad, G[tag + "/dx"], 1e-6, tag + " dx")
    if tag == "scatter16":  # empty slot (bbox -1): exact zeros
        assert float(y[1].detach().abs().max()) == 0.0


@pytest.mark.parametrize("tag", ["b3", "b4"])
def test_global_attention(tag):
    G, _ = gu.load("attention_cases")
    h = gu.full(G, tag + "/h").requires_grad_(True)
    ctx = gu.full(G, tag + "/ctx").requires_grad_(True)
    w = gu.full(G, tag + "/w").requires_grad_(True)
    mask = gu.full(G, tag + "/mask").bool()
    wc, attn = O.global_attention(h, ctx, mask, {"a.conv_context.weight": w}, "a")
    gu.check(wc, G[tag + "/wc"], 1e-6)
    gu.check(attn, G[tag + "/attn"], 1e-6)
    wc.backward(gu.full(G, tag + "/g"))
    gu.check(h.grad, G[tag + "/dh"], 1e-5)
    gu.check(ctx.grad, G[tag + "/dctx"], 1e-5)
    gu.check(w.grad, G[tag + "/dw"], 1e-5)


def test_damsm_losses():
    G, _ = gu.load("attention_cases")
    cfg = O.Cfg()
    feat = gu.full(G, "damsm/feat").requires_grad_(True)
    code = gu.full(G, "damsm/code").requires_grad_(True)
    words, sent = gu.full(G, "damsm/words"), gu.full(G, "damsm/sent")
    lens = gu.full(G, "damsm/lens").long()
    B = feat.shape[0]
    labels = torch.arange(B)
    w0, w1 = O.words_loss(feat, words, labels, lens, np.arange(B), B, cfg)
    s0, s1 = O.sent_loss(code, sent, labels, np.arange(B), B, cfg)
    for k, v in {"w0": w0, "w1": w1, "s0": s0, "s1": s1}.items():
        gu.check(v, G["damsm/" + k], 1e-5, k)
    (w0 + w1 + s0 + s1).backward()
    gu.check(feat.grad, G["damsm/dfeat"], 1e-5)
    gu.check(code.grad, G["damsm/dcode"], 1e-5)


def test_attngan_step_matches_reference(keys):
    """G forward, 3 D losses + grads, full G loss (adversarial + DAMSM via stand-in + KL) + grads,
    BatchNorm running statistics -- all against the executed reference."""
    G, meta = gu.load("attngan_tiny_step")
    c, seed = meta["cfg"], meta["seed"]
    cfg = _tiny_cfg(c)
    batch = synth.attngan_batch(c["B"], T=c["T"], nef=c["EMBEDDING_DIM"], nz=c["Z_DIM"], seed=seed)
    PG = _build(keys["tiny"]["G_NET"], seed + 1)
    PDs = [_build(keys["tiny"]["D_NET%d" % (64 << i)], seed + 2 + i) for i in range(3)]
    eps = gu.full(G, "G/eps")
    B = c["B"]
    real, fake = torch.ones(B), torch.zeros(B)
    tm, tmi, oh = batch["transf_matrices"], batch["transf_matrices_inv"], batch["label_one_hot"]
    imgs, atts, mu, logvar = O.g_net(PG, cfg, batch["noise"], batch["sent_emb"], batch["words_embs"],
                                     batch["mask"], tmi, oh, eps=eps)
    for i in range(3):
        gu.check(imgs[i], G["G/fake%d" % i], TOL, "fake%d" % i)
    for i in range(2):
        gu.check(atts[i], G["G/att%d" % i], TOL, "att%d" % i)
    gu.check(mu, G["G/mu"], TOL)
    gu.check(logvar, G["G/logvar"], TOL)
    for i, PD in enumerate(PDs):
        kw = dict(label=oh, theta=tm, theta_inv=tmi) if i == 0 else {}
        errD = O.discriminator_loss(i, PD, cfg, batch["imgs"][i], imgs[i], batch["sent_emb"], real, fake, **kw)
        gu.check(errD, G["D%d/errD" % i], TOL, "errD%d" % i)
        names = [k for k, p in PD.items() if p.requires_grad]
        grads = torch.autograd.grad(errD, [PD[k] for k in names])
        for k, g in zip(names, grads):
            gu.check(g, G["D%d/grad/%s" % (i, k)], 5e-5, "D%d grad %s" % (i, k))
        for k, v in PD.items():
            if "running" in k:
                gu.check(v, G["D%d/buf_after_dstep/%s" % (i, k)], TOL, k)
    enc = synth.StandInEncoder(c["EMBEDDING_DIM"])
    adv = O.generator_gan_loss(PDs, cfg, imgs, batch["sent_emb"], real, oh, tm, tmi)
    feat, code = enc(imgs[-1])
    labels = torch.arange(B)
    w0, w1 = O.words_loss(feat, batch["words_embs"], labels, batch["cap_lens"], batch["class_ids"], B, cfg)
    s0, s1 = O.sent_loss(code, batch["sent_emb"], labels, batch["class_ids"], B, cfg)
    total = adv + (w0 + w1) * cfg.LAMBDA + (s0 + s1) * cfg.LAMBDA
    kl = O.kl_loss(mu, logvar)
    gu.check(total, G["G/errG_total"], TOL, "errG_total")
    gu.check(kl, G["G/kl"], TOL, "kl")
    names = [k for k, p in PG.items() if p.requires_grad]
    grads = torch.autograd.grad(total + kl, [PG[k] for k in names])
    for k, g in zip(names, grads):
        gu.check(g, G["G/grad/%s" % k], 2e-4, "G grad %s" % k)
    for k, v in PG.items():
        if "running" in k:
            gu.check(v, G["G/buf/%s" % k], TOL, k)


def test_attngan_gd_only_step(keys):
    """oracle.gd_step (the G+D-only step the bench times) against the executed reference."""
    G, meta = gu.load("attngan_tiny_gd")
    c, seed = meta["cfg"], meta["seed"]
    cfg = _tiny_cfg(c)
    batch = synth.attngan_batch(c["B"], T=c["T"], nef=c["EMBEDDING_DIM"], nz=c["Z_DIM"], seed=seed)
    PG = _build(keys["tiny"]["G_NET"], seed + 1)
    PDs = [_build(keys["tiny"]["D_NET%d" % (64 << i)], seed + 2 + i) for i in range(3)]
    out = O.gd_step(PG, PDs, cfg, batch, eps=gu.full(G, "G/eps"))
    gu.check(out["errG"], G["G/errG_adv"], TOL)
    gu.check(out["kl"], G["G/kl"], TOL)
    for k, p in PG.items():
        if p.requires_grad:
            gu.check(p.grad, G["G/grad/%s" % k], 2e-4, "G grad %s" % k)


def test_train_steps_with_adam_and_ema(keys):
    """a20: three consecutive iterations of trainer.py:294-342 INCLUDING the optimiser steps (each discriminator is
    updated before the generator step sees it), the DAMSM branch and the EMA: the oracle's ``train_step`` against the
    executed reference (tests/golden/make_golden_trainstep.py)."""
    G, meta = gu.load("attngan_tiny_trainstep")
    c, seed, K = meta["cfg"], meta["seed"], meta["steps"]
    cfg = _tiny_cfg(c)
    PG = _build(keys["tiny"]["G_NET"], seed + 1)
    PDs = []
    for i in range(3):
        shapes = {k: torch.empty(s) for k, s in keys["tiny"]["D_NET%d" % (64 << i)].items()}
        PDs.append(O.leafify(synth.soften_logits(synth.fill_state_dict(shapes, seed + 2 + i), meta["logit_scale"])))
    batch = synth.attngan_batch(c["B"], T=c["T"], nef=c["EMBEDDING_DIM"], nz=c["Z_DIM"], seed=seed)
    enc = synth.StandInEncoder(c["EMBEDDING_DIM"])
    state = O.make_train_state(PG, PDs)
    for k in range(K):
        noise = torch.from_numpy(np.random.RandomState(seed + 10 + k).standard_normal((c["B"], c["Z_DIM"])).astype(np.float32))
        out = O.train_step(PG, PDs, state, cfg, batch, eps=gu.full(G, "step%d/eps" % k), noise=noise, image_encoder=enc)
        gu.check(sum(out["errD"]), G["step%d/errD_total" % k], 1e-4, "errD step %d" % k)
        gu.check(out["errG"] + out["kl"], G["step%d/errG_total" % k], 1e-4, "errG step %d" % k)
        gu.check(out["kl"], G["step%d/kl" % k], 1e-4, "kl step %d" % k)
    check_train_state(G, {"G": (PG, state["optG"].state), "D0": (PDs[0], state["optDs"][0].state),
                          "D1": (PDs[1], state["optDs"][1].state), "D2": (PDs[2], state["optDs"][2].state)},
                      dict(zip([n for n, p in PG.items() if p.requires_grad], state["ema"])))


# Gates of the after-K-steps comparison.  One Adam step moves every weight by ~lr = 2e-4 whatever the size of its
# gradient (m / sqrt(v) = +-1 at t = 1), so summation-order noise on near-zero gradient entries shows up at full size in
# those elements: fp32-vs-fp32 restatements of the same step differ by up to 1.2e-4 per parameter tensor (5e-5 per network)
# and up to 2e-2 in the moments of single tensors whose gradient nearly cancels (6e-4 per network; measured: oracle vs
# reference, both CPU fp32; the GPU fp32 path, whose summation order differs, reaches 6e-4 on h_net1.bbox_net.encode.0.weight --
# one-hot label planes, gradient entries spanning 7 decades -- and 1.1e-3 on the running mean of the BatchNorm behind it).
# The tiny nets are also ill-conditioned: perturbing the discriminator parameters by 2e-4 (one Adam step of sign noise) changes
# the ORACLE's own gradients by 1-2e-3 over a whole network and up to 5e-2 on single tensors (measured), so from the second
# step on the moments cannot agree better than a few 1e-2.
# D_NET256 is the worst: a 3e-6 perturbation of the generated 256^2 image -- the rounding difference between two fp32
# implementations of G -- changes its fp64 gradients by up to 2.5e-3 (img_code_s16.3.bias; measured on the CPU oracle).
# The moments are gated per NETWORK (norm-weighted); the per-tensor gate is a sanity bound only (tensors whose gradient nearly
# cancels -- img_net2.img.0.weight, the label BatchNorm bias -- sit at 0.02-0.13 from rounding alone).
# A missing or doubled update would be >= 6e-3 on the parameters, O(1) on the moments and >= 0.1 on a running statistic.
TS_PARAM_RMS, TS_PARAM_NET, TS_MOM_NET, TS_MOM, TS_EMA, TS_BUF = 1e-4, 5e-4, 5e-2, 0.5, 1e-5, 2e-3


def check_train_state(G, nets, ema, scale=1.0):
    """nets: tag -> (name -> tensor incl. buffers, param tensor -> Adam state); ema: G param name -> EMA tensor."""
    for tag, (P, ostate) in nets.items():
        ap, am, av = gu.Aggregate(), gu.Aggregate(), gu.Aggregate()
        for name, p in P.items():
            if ("%s/param/%s" % (tag, name)) in G:
                e = ap.add(p, G["%s/param/%s" % (tag, name)], name)
                # per tensor: rms deviation below half an Adam step (lr = 2e-4 moves every entry by ~lr per step; a skipped or
                # doubled step would show as >= lr).  Small-valued tensors (BatchNorm beta ~ 0.02) make rel-L2 meaningless here.
                summ = G["%s/param/%s" % (tag, name)]
                ref = np.asarray(summ["full"] if "full" in summ else summ["sample"], np.float64)
                a = p.detach().cpu().double().numpy().reshape(-1)
                got = a if "full" in summ else a[gu._idx(a.size)]
                rms = float(np.sqrt(np.mean((got - ref) ** 2)))
                assert rms <= scale * TS_PARAM_RMS, "%s param %s: rms deviation %.3e (rel-L2 %.3e)" % (tag, name, rms, e)
                s = ostate[p]
                e = am.add(s["exp_avg"], G["%s/exp_avg/%s" % (tag, name)], name)
                assert e <= scale * TS_MOM, "%s exp_avg %s: %.3e" % (tag, name, e)
                e = av.add(s["exp_avg_sq"], G["%s/exp_avg_sq/%s" % (tag, name)], name)
                assert e <= scale * TS_MOM, "%s exp_avg_sq %s: %.3e" % (tag, name, e)
            elif ("%s/buf/%s" % (tag, name)) in G:
                gu.check(p.float(), G["%s/buf/%s" % (tag, name)], scale * TS_BUF, "%s buffer %s" % (tag, name))
        assert ap.rel() <= scale * TS_PARAM_NET, "%s params: %.3e (worst %s)" % (tag, ap.rel(), ap.worst)
        assert am.rel() <= scale * TS_MOM_NET, "%s exp_avg: %.3e (worst %s)" % (tag, am.rel(), am.worst)
        assert av.rel() <= scale * TS_MOM_NET, "%s exp_avg_sq: %.3e (worst %s)" % (tag, av.rel(), av.worst)
        print(tag, "params %.2e (worst %.2e)  exp_avg %.2e (worst %.2e)  exp_avg_sq %.2e (worst %.2e)"
              % (ap.rel(), ap.worst[0], am.rel(), am.worst[0], av.rel(), av.worst[0]))
    for name, a in ema.items():
        gu.check(a, G["G/ema/%s" % name], TS_EMA, "ema %s" % name)
