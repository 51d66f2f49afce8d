"""Parity where the benchmark runs: BASELINE.json config 5 dimensions (GF 48, DF 96, T 18, R_NUM 3, nef 256), product
precision (bf16x3), the FULL step of trainer.py:294-340 -- G forward, three discriminator losses + backward, generator loss
WITH the DAMSM words / sentence branch through the libmog ``CNN_ENCODER`` (Inception-v3), KL, backward -- against the CPU
oracle (``oracle.attngan_oracle.gd_step`` + ``oracle.encoder_oracle``, both pinned on the executed reference).  B = 4 keeps
the oracle at ~10 s; the layer shapes (96/192-channel 128^2 / 256^2 halo tiles, K = 24576 / 27648 split-K layers,
two-segment discriminator passes) are the benchmark's.  Run on the B200 box: -m gpu."""
import numpy as np
import pytest
import torch

import golden_util as gu
from mog_b200 import synth
from oracle import attngan_oracle as O

pytestmark = pytest.mark.gpu

import os
C5 = dict(GF_DIM=48, DF_DIM=96, Z_DIM=100, R_NUM=3, EMBEDDING_DIM=256, T=18, B=int(os.environ.get("C5_B", "4")))
C5_PREC = os.environ.get("C5_PREC", "bf16x3")
# Stated tolerances (rel-L2 against the fp32 CPU oracle).  Outputs: images and losses <= 2e-4 in the product precision
# (north star: 1e-3).  Gradients: norm-weighted rel-L2 over ALL parameters of a network; discriminators <= 2e-3.  The
# generator's gradient passes backwards through three discriminators with batch-statistics BatchNorm over 2-4 samples and
# is ill-conditioned in this synthetic setting -- the EXACT fp32 CUDA-core mode already differs from the CPU oracle by
# 4e-3 (B=4) to 8e-3 (B=2) without and 9e-3 to 2e-2 with the DAMSM branch (measured) -- so the product precision is gated
# against that floor: bf16x3 (operands carry 16 mantissa bits: ~5e-6 per conv against ~3e-7) must stay within
# max(2e-3, 5 x the error of the exact mode on the same inputs); measured at B=4: 1.3e-2 vs 3.6e-3, and 3.1e-2 vs 8.9e-3 with
# the DAMSM branch.
IMG_TOL, LOSS_TOL = 2e-4, 2e-4
GRAD_TOL = 2e-3
G_FLOOR_TOL = 5e-2       # exact-mode sanity bound on the generator gradient (all parameters)
GRAD_TOL_TENSOR = 0.15   # sanity bound per tensor


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


def _g_forward(s):
    b = s["b"]
    return s["netG"](b["noise"], b["sent_emb"], b["words_embs"], b["mask"], b["transf_matrices_inv"], b["label_one_hot"], eps=b["eps"])


def _agg(pairs):
    num = sum(float((a.detach().cpu().double() - r.double()).pow(2).sum()) for a, r in pairs)
    den = sum(float(r.double().pow(2).sum()) for a, r in pairs)
    return (num / den) ** 0.5


def _d_steps(s, prec):
    from mog_b200 import ops
    from mog_b200.attngan.miscc import losses as L
    ops.set_precision(prec)
    b, ref, B = s["b"], s["ref_full"], s["c"]["B"]
    imgs, _, mu, logvar = _g_forward(s)
    out = {"img": [_rel(imgs[i], ref["fake_imgs"][i]) for i in range(3)], "loss": [], "agg": [], "worst": (0.0, "")}
    real, fake = torch.ones(B, device="cuda"), torch.zeros(B, device="cuda")
    for i, netD in enumerate(s["netsD"]):
        netD.zero_grad(set_to_none=True)
        kw = dict(local_labels=b["label_one_hot"], transf_matrices=b["transf_matrices"],
                  transf_matrices_inv=b["transf_matrices_inv"]) if i == 0 else {}
        errD = L.discriminator_loss(netD, b["imgs"][i], imgs[i], b["sent_emb"], real, fake, [0], **kw)
        errD.backward()
        out["loss"].append(abs(float(errD) - float(ref["errD"][i])) / abs(float(ref["errD"][i])))
        pairs = [(p.grad, ref["Dgrad"][i][k]) for k, p in netD.named_parameters()]
        out["agg"].append(_agg(pairs))
        for k, p in netD.named_parameters():
            out["worst"] = max(out["worst"], (_rel(p.grad, ref["Dgrad"][i][k]), "D%d %s" % (i, k)))
    return out


def test_config5_g_forward_and_d_steps(setup):
    """G forward (3 images) + the three discriminator steps (loss, all parameter gradients) at config-5 widths."""
    x3, fp = _d_steps(setup, "bf16x3"), _d_steps(setup, "fp32")
    print("config 5 bf16x3: images %s  losses %s  D grads (all-parameter) %s  worst tensor %.2e %s" %
          (["%.1e" % e for e in x3["img"]], ["%.1e" % e for e in x3["loss"]], ["%.1e" % e for e in x3["agg"]], *x3["worst"]))
    print("config 5 fp32  : images %s  losses %s  D grads (all-parameter) %s  worst tensor %.2e %s" %
          (["%.1e" % e for e in fp["img"]], ["%.1e" % e for e in fp["loss"]], ["%.1e" % e for e in fp["agg"]], *fp["worst"]))
    for r in (x3, fp):
        assert max(r["img"]) < IMG_TOL and max(r["loss"]) < LOSS_TOL
        assert max(r["agg"]) < GRAD_TOL
        assert r["worst"][0] < GRAD_TOL_TENSOR


def _g_step(s, prec, damsm):
    from mog_b200 import ops
    from mog_b200.attngan.miscc import losses as L
    ops.set_precision(prec)
    b, B = s["b"], s["c"]["B"]
    ref = s["ref_full"] if damsm else s["ref_gd"]
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
        named = list(s["netG"].named_parameters())
        worst = max((_rel(p.grad, ref["Ggrad"][k]), k) for k, p in named)
        return {"errG": abs(float(errG) - float(ref["errG"])) / abs(float(ref["errG"])),
                "kl": abs(float(kl) - float(ref["kl"])) / abs(float(ref["kl"])),
                "agg": _agg([(p.grad, ref["Ggrad"][k]) for k, p in named]), "worst": worst}
    finally:
        for d in s["netsD"]:
            for p in d.parameters():
                p.requires_grad_(True)


@pytest.mark.parametrize("damsm", [False, True])
def test_config5_generator_step(setup, damsm):
    """generator_loss (losses.py:177-226) without / with the ranking branch through the libmog CNN_ENCODER (Inception-v3 +
    words / sentence losses), + KL, all generator gradients, at config-5 widths."""
    x3, fp = _g_step(setup, "bf16x3", damsm), _g_step(setup, "fp32", damsm)
    for name, r in (("bf16x3", x3), ("fp32  ", fp)):
        print("config 5 G step (damsm=%s) %s: errG %.1e kl %.1e  G grads all-parameter %.2e  worst tensor %.2e %s"
              % (damsm, name, r["errG"], r["kl"], r["agg"], *r["worst"]))
        assert r["errG"] < LOSS_TOL and r["kl"] < LOSS_TOL
        assert r["worst"][0] < GRAD_TOL_TENSOR
    assert fp["agg"] < G_FLOOR_TOL
    assert x3["agg"] < max(GRAD_TOL, 5.0 * fp["agg"]), (x3["agg"], fp["agg"])
