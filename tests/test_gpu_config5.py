"""Parity where the benchmark runs: BASELINE.json config 5 dimensions (GF 48, DF 96, T 18, R_NUM 3, nef 256), product
precision (bf16x3), the FULL step of trainer.py:294-340 -- G forward, three discriminator losses + backward, generator loss
WITH the DAMSM words / sentence branch through the libmog ``CNN_ENCODER`` (Inception-v3), KL, backward -- against the CPU
oracle (``oracle.attngan_oracle.gd_step`` + ``oracle.encoder_oracle``, both pinned on the executed reference).  B = 2 keeps
the oracle at a few seconds; the layer shapes (96/192-channel 128^2 / 256^2 halo tiles, K = 24576 / 27648 split-K layers,
two-segment discriminator passes) are the benchmark's.  Run on the B200 box: -m gpu."""
import numpy as np
import pytest
import torch

import golden_util as gu
from mog_b200 import synth
from oracle import attngan_oracle as O

pytestmark = pytest.mark.gpu

C5 = dict(GF_DIM=48, DF_DIM=96, Z_DIM=100, R_NUM=3, EMBEDDING_DIM=256, T=18, B=2)
# stated tolerances (rel-L2 against the fp32 CPU oracle), product precision bf16x3
IMG_TOL, LOSS_TOL = 2e-4, 2e-4
GRAD_TOL = 2e-3          # every parameter gradient of the three discriminators and of the generator's adversarial+KL step
GRAD_TOL_DAMSM = 3e-2    # generator gradients of the full loss: they pass through ~95 ReLU layers of the frozen Inception-v3


def _rel(a, b):
    return gu.rel_l2(a.detach().cpu().numpy(), b.detach().cpu().numpy())


@pytest.fixture(scope="module")
def setup():
    from mog_b200 import ops
    from mog_b200.attngan import model as M
    from mog_b200.attngan.miscc.config import cfg, reset_cfg
    reset_cfg()
    c = C5
    cfg.GAN.GF_DIM, cfg.GAN.DF_DIM, cfg.GAN.Z_DIM, cfg.GAN.R_NUM = c["GF_DIM"], c["DF_DIM"], c["Z_DIM"], c["R_NUM"]
    cfg.TEXT.EMBEDDING_DIM, cfg.TEXT.WORDS_NUM = c["EMBEDDING_DIM"], c["T"]
    cfg.TRAIN.SMOOTH.GAMMA1, cfg.TRAIN.SMOOTH.GAMMA2, cfg.TRAIN.SMOOTH.GAMMA3, cfg.TRAIN.SMOOTH.LAMBDA = 4.0, 5.0, 10.0, 50.0
    cfg.TRAIN.BATCH_SIZE = c["B"]
    old = ops.get_precision()
    ops.set_precision("bf16x3")
    seed = 500
    netG = M.G_NET()
    netsD = [M.D_NET64(), M.D_NET128(), M.D_NET256()]
    netG.load_state_dict(synth.fill_state_dict(netG.state_dict(), seed + 1))
    for i, d in enumerate(netsD):
        d.load_state_dict(synth.soften_logits(synth.fill_state_dict(d.state_dict(), seed + 2 + i), 0.02))
    enc = M.CNN_ENCODER(c["EMBEDDING_DIM"])
    enc.load_state_dict(synth.fill_encoder_state_dict(enc.state_dict(), 9))
    PE = {k: v.detach().clone() for k, v in enc.state_dict().items()}
    for p in enc.parameters():
        p.requires_grad = False
    PG = O.leafify(netG.state_dict())
    PDs = [O.leafify(d.state_dict()) for d in netsD]
    batch = synth.attngan_batch(c["B"], T=c["T"], nef=c["EMBEDDING_DIM"], nz=c["Z_DIM"], seed=seed)
    ocfg = O.Cfg(**{k: v for k, v in c.items() if k not in ("T", "B")})
    torch.set_num_threads(max(1, torch.get_num_threads()))
    ref_full = O.gd_step(PG, PDs, ocfg, batch, PE=PE)
    ref_full["Ggrad"] = {k: p.grad.clone() for k, p in PG.items() if p.requires_grad and p.grad is not None}
    ref_full["Dgrad"] = [{k: p.grad.clone() for k, p in PD.items() if p.requires_grad} for PD in PDs]
    ref_gd = O.gd_step(PG, PDs, ocfg, batch)              # adversarial + KL only (no DAMSM)
    ref_gd["Ggrad"] = {k: p.grad.clone() for k, p in PG.items() if p.requires_grad and p.grad is not None}
    netG.cuda().train()
    for d in netsD:
        d.cuda().train()
    enc.cuda().eval()
    b = {}
    for k, v in batch.items():
        b[k] = [t.cuda() for t in v] if isinstance(v, list) else (v.cuda() if torch.is_tensor(v) else v)
    yield dict(netG=netG, netsD=netsD, enc=enc, b=b, ref_full=ref_full, ref_gd=ref_gd, c=c)
    ops.set_precision(old)


def _g_forward(s):
    b = s["b"]
    return s["netG"](b["noise"], b["sent_emb"], b["words_embs"], b["mask"], b["transf_matrices_inv"], b["label_one_hot"], eps=b["eps"])


def test_config5_g_forward_and_d_steps(setup):
    from mog_b200.attngan.miscc import losses as L
    s, b, ref = setup, setup["b"], setup["ref_full"]
    B = s["c"]["B"]
    imgs, _, mu, logvar = _g_forward(s)
    for i in range(3):
        e = _rel(imgs[i], ref["fake_imgs"][i])
        assert e < IMG_TOL, "fake image %d: %.3e" % (i, e)
    real, fake = torch.ones(B, device="cuda"), torch.zeros(B, device="cuda")
    worst = (0.0, "")
    for i, netD in enumerate(s["netsD"]):
        netD.zero_grad(set_to_none=True)
        kw = dict(local_labels=b["label_one_hot"], transf_matrices=b["transf_matrices"],
                  transf_matrices_inv=b["transf_matrices_inv"]) if i == 0 else {}
        errD = L.discriminator_loss(netD, b["imgs"][i], imgs[i], b["sent_emb"], real, fake, [0], **kw)
        errD.backward()
        assert abs(float(errD) - float(ref["errD"][i])) <= LOSS_TOL * abs(float(ref["errD"][i])), (i, float(errD), float(ref["errD"][i]))
        for k, p in netD.named_parameters():
            e = _rel(p.grad, ref["Dgrad"][i][k])
            worst = max(worst, (e, "D%d %s" % (i, k)))
            assert e < GRAD_TOL, "D%d grad %s: %.3e" % (i, k, e)
    print("config 5 D gradients: worst rel-L2 %.2e (%s)" % worst)


@pytest.mark.parametrize("damsm", [False, True])
def test_config5_generator_step(setup, damsm):
    """generator_loss (losses.py:177-226) without / with the ranking branch through libmog CNN_ENCODER, + KL, all G grads."""
    from mog_b200.attngan.miscc import losses as L
    s, b = setup, setup["b"]
    ref = s["ref_full"] if damsm else s["ref_gd"]
    B = s["c"]["B"]
    imgs, _, mu, logvar = _g_forward(s)
    for d in s["netsD"]:
        for p in d.parameters():
            p.requires_grad_(False)
    try:
        s["netG"].zero_grad(set_to_none=True)
        real = torch.ones(B, device="cuda")
        match = torch.arange(B, device="cuda")
        errG, _ = L.generator_loss(s["netsD"], s["enc"] if damsm else None, imgs, real, b["words_embs"], b["sent_emb"], match,
                                   b["cap_lens"], b["class_ids"], [0], local_labels=b["label_one_hot"],
                                   transf_matrices=b["transf_matrices"], transf_matrices_inv=b["transf_matrices_inv"])
        kl = L.KL_loss(mu, logvar)
        (errG + kl).backward()
        assert abs(float(errG) - float(ref["errG"])) <= LOSS_TOL * abs(float(ref["errG"])), (float(errG), float(ref["errG"]))
        assert abs(float(kl) - float(ref["kl"])) <= LOSS_TOL * abs(float(ref["kl"]))
        tol = GRAD_TOL_DAMSM if damsm else GRAD_TOL
        worst = (0.0, "")
        agg_n = agg_d = 0.0
        for k, p in s["netG"].named_parameters():
            r = ref["Ggrad"][k]
            e = _rel(p.grad, r)
            worst = max(worst, (e, k))
            agg_n += float((p.grad.detach().cpu().double() - r.double()).pow(2).sum())
            agg_d += float(r.double().pow(2).sum())
            assert e < tol, "G grad %s (damsm=%s): %.3e" % (k, damsm, e)
        print("config 5 G gradients (damsm=%s): all-parameter rel-L2 %.2e, worst tensor %.2e (%s)"
              % (damsm, (agg_n / agg_d) ** 0.5, worst[0], worst[1]))
    finally:
        for d in s["netsD"]:
            for p in d.parameters():
                p.requires_grad_(True)
