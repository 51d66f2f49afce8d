#!/usr/bin/env python
"""Generate the golden fixtures by EXECUTING THE UNMODIFIED REFERENCE in this container.

    python tests/golden/make_golden.py            # writes tests/golden/*.npz, *.json

Imports ``/root/reference/code/coco/attngan/{model,GlobalAttention,miscc/losses,miscc/utils}.py``
through the harness shims of SURVEY.md section 8(c) (easydict / cPickle / skimage stand-ins, CPU aliases
for ``torch.cuda.FloatTensor``, pass-through ``data_parallel``, ``torch.ByteTensor`` -> bool) and
runs them on the deterministic synthetic inputs / weights of ``mog_b200.synth``.  The reference
cannot travel to the GPU box, so only the resulting vectors are committed; the oracle
(``oracle/attngan_oracle.py``) and the CUDA path are both checked against them.
"""
import json
import os
import sys

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/code/coco/attngan"
sys.path[:0] = [os.path.join(ROOT, "baseline", "shims"), REF, os.path.join(ROOT, "multiple-objects-gan_b200"),
                os.path.join(ROOT, "tests")]

# ---- shims; CPU-only aliases ------------------------------------------------
torch.cuda.FloatTensor = torch.FloatTensor
torch.cuda.DoubleTensor = torch.DoubleTensor


def _dp(module, inputs, device_ids=None, **kw):
    return module(*inputs) if isinstance(inputs, tuple) else module(inputs)


nn.parallel.data_parallel = _dp
torch.ByteTensor = lambda a: torch.as_tensor(a).bool()

from miscc.config import cfg  # noqa: E402  (reference)
import model as M  # noqa: E402  (reference, unmodified)
import GlobalAttention as GA  # noqa: E402
from miscc import losses as L  # noqa: E402
from miscc import utils as U  # noqa: E402

from mog_b200 import synth  # noqa: E402
from golden_util import summarize  # noqa: E402


def set_cfg(c):
    cfg.CUDA = False
    cfg.GAN.GF_DIM, cfg.GAN.DF_DIM, cfg.GAN.Z_DIM = c["GF_DIM"], c["DF_DIM"], c["Z_DIM"]
    cfg.GAN.CONDITION_DIM, cfg.GAN.R_NUM = 100, c["R_NUM"]
    cfg.TEXT.EMBEDDING_DIM, cfg.TEXT.WORDS_NUM = c["EMBEDDING_DIM"], c["T"]
    cfg.TREE.BRANCH_NUM = 3
    cfg.TRAIN.BATCH_SIZE = c["B"]
    cfg.TRAIN.SMOOTH.GAMMA1, cfg.TRAIN.SMOOTH.GAMMA2 = 4.0, 5.0
    cfg.TRAIN.SMOOTH.GAMMA3, cfg.TRAIN.SMOOTH.LAMBDA = 10.0, 50.0


def flat(prefix, d, out):
    for k, v in d.items():
        out[prefix + "/" + k] = v


def save(name, entries, meta):
    arrays = {}
    index = {}
    for k, s in entries.items():
        index[k] = {"shape": s["shape"], "l2": s["l2"], "sum": s["sum"],
                    "kind": "full" if "full" in s else "sample"}
        arrays[k] = s["full"] if "full" in s else s["sample"]
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **arrays)
    with open(os.path.join(HERE, name + ".json"), "w") as f:
        json.dump({"meta": meta, "index": index}, f, indent=1, sort_keys=True)
    print("wrote", name, "entries", len(entries),
          "bytes", os.path.getsize(os.path.join(HERE, name + ".npz")))


def grads_of(net):
    return {k: p.grad for k, p in net.named_parameters() if p.grad is not None}


def attngan_step(name, c, seed):
    """trainer.py:294-340 on synthetic data (no optimiser step; the DAMSM image encoder is the
    fixed stand-in of mog_b200.synth.StandInEncoder, passed through generator_loss' own
    ``image_encoder`` argument)."""
    set_cfg(c)
    B, T, nef = c["B"], c["T"], c["EMBEDDING_DIM"]
    batch = synth.attngan_batch(B, T=T, nef=nef, nz=c["Z_DIM"], seed=seed)
    netG = M.G_NET()
    netsD = [M.D_NET64(), M.D_NET128(), M.D_NET256()]
    netG.load_state_dict(synth.fill_state_dict(netG.state_dict(), seed + 1))
    for i, d in enumerate(netsD):
        d.load_state_dict(synth.fill_state_dict(d.state_dict(), seed + 2 + i))
    for n in [netG] + netsD:
        n.train()
    keys = {"G_NET": {k: list(v.shape) for k, v in netG.state_dict().items()}}
    for i, d in enumerate(netsD):
        keys["D_NET%d" % (64 << i)] = {k: list(v.shape) for k, v in d.state_dict().items()}

    # the reference draws eps from the global CPU generator inside CA_NET (model.py:338): seed,
    # record the draw, re-seed so the forward consumes the very same numbers.
    torch.manual_seed(seed)
    eps = torch.FloatTensor(B, 100).normal_()
    torch.manual_seed(seed)

    E = {}
    real_labels, fake_labels = torch.ones(B), torch.zeros(B)
    match_labels = torch.arange(B)
    tm, tmi, onehot = batch["transf_matrices"], batch["transf_matrices_inv"], batch["label_one_hot"]
    fake_imgs, att_maps, mu, logvar = netG(batch["noise"], batch["sent_emb"], batch["words_embs"],
                                           batch["mask"], tmi, onehot)
    for i, f in enumerate(fake_imgs):
        E["G/fake%d" % i] = summarize(f)
    for i, a in enumerate(att_maps):
        E["G/att%d" % i] = summarize(a)
    E["G/mu"], E["G/logvar"], E["G/eps"] = summarize(mu), summarize(logvar), summarize(eps)
    E["G/eps"]["full"] = eps.numpy().reshape(-1)

    for i, netD in enumerate(netsD):
        netD.zero_grad()
        if i == 0:
            errD = L.discriminator_loss(netD, batch["imgs"][i], fake_imgs[i], batch["sent_emb"],
                                        real_labels, fake_labels, [0], local_labels=onehot,
                                        transf_matrices=tm, transf_matrices_inv=tmi)
        else:
            errD = L.discriminator_loss(netD, batch["imgs"][i], fake_imgs[i], batch["sent_emb"],
                                        real_labels, fake_labels, [0])
        errD.backward()
        E["D%d/errD" % i] = summarize(errD)
        for k, g in grads_of(netD).items():
            E["D%d/grad/%s" % (i, k)] = summarize(g)
        for k, v in netD.state_dict().items():
            if "running" in k:
                E["D%d/buf_after_dstep/%s" % (i, k)] = summarize(v)

    # features of the real images, recomputed on a scratch copy (buffers untouched above)
    netG.zero_grad()
    enc = synth.StandInEncoder(nef)
    errG_total, _ = L.generator_loss(netsD, enc, fake_imgs, real_labels, batch["words_embs"],
                                     batch["sent_emb"], match_labels, batch["cap_lens"],
                                     batch["class_ids"], [0], local_labels=onehot,
                                     transf_matrices=tm, transf_matrices_inv=tmi)
    kl = L.KL_loss(mu, logvar)
    E["G/errG_total"], E["G/kl"] = summarize(errG_total), summarize(kl)
    (errG_total + kl).backward()
    for k, g in grads_of(netG).items():
        E["G/grad/%s" % k] = summarize(g)
    for k, v in netG.state_dict().items():
        if "running" in k:
            E["G/buf/%s" % k] = summarize(v)
    save(name, E, {"cfg": c, "seed": seed, "what": "reference attngan G fwd + 3 D steps + G step "
                   "(DAMSM via StandInEncoder), no optimiser"})
    return keys


def attngan_gd_only(name, c, seed):
    """Same as above but the G loss is the adversarial part + KL only (no DAMSM branch): the
    reference loop of losses.py:189-204 restated with the reference's own nets and heads."""
    set_cfg(c)
    B, T, nef = c["B"], c["T"], c["EMBEDDING_DIM"]
    batch = synth.attngan_batch(B, T=T, nef=nef, nz=c["Z_DIM"], seed=seed)
    netG = M.G_NET()
    netsD = [M.D_NET64(), M.D_NET128(), M.D_NET256()]
    netG.load_state_dict(synth.fill_state_dict(netG.state_dict(), seed + 1))
    for i, d in enumerate(netsD):
        d.load_state_dict(synth.fill_state_dict(d.state_dict(), seed + 2 + i))
    torch.manual_seed(seed)
    eps = torch.FloatTensor(B, 100).normal_()
    torch.manual_seed(seed)
    E = {}
    real_labels = torch.ones(B)
    tm, tmi, onehot = batch["transf_matrices"], batch["transf_matrices_inv"], batch["label_one_hot"]
    fake_imgs, _, mu, logvar = netG(batch["noise"], batch["sent_emb"], batch["words_embs"],
                                    batch["mask"], tmi, onehot)
    E["G/eps"] = summarize(eps)
    E["G/eps"]["full"] = eps.numpy().reshape(-1)
    total = 0
    bce = nn.BCELoss()
    for i, netD in enumerate(netsD):
        f = netD(fake_imgs[i], onehot, tm, tmi) if i == 0 else netD(fake_imgs[i])
        total = total + bce(netD.UNCOND_DNET(f), real_labels) + bce(netD.COND_DNET(f, batch["sent_emb"]), real_labels)
    kl = L.KL_loss(mu, logvar)
    E["G/errG_adv"], E["G/kl"] = summarize(total), summarize(kl)
    netG.zero_grad()
    (total + kl).backward()
    for k, g in grads_of(netG).items():
        E["G/grad/%s" % k] = summarize(g)
    save(name, E, {"cfg": c, "seed": seed, "what": "reference attngan G adversarial+KL loss and grads"})


def attention_cases(name):
    """GlobalAttentionGeneral (incl. the mask-tiling quirk for B not dividing queryL) and
    func_attention / words_loss / sent_loss on synthetic features."""
    E = {}
    rng = np.random.RandomState(5)
    for tag, (B, idf, cdf, ih, T) in {"b3": (3, 8, 16, 8, 5), "b4": (4, 48, 32, 16, 18)}.items():
        att = GA.GlobalAttentionGeneral(idf, cdf)
        w = torch.from_numpy(rng.standard_normal((idf, cdf, 1, 1)).astype(np.float32) / np.sqrt(cdf))
        att.conv_context.weight.data.copy_(w)
        h = torch.from_numpy(rng.standard_normal((B, idf, ih, ih)).astype(np.float32)).requires_grad_(True)
        ctx = torch.from_numpy(np.tanh(rng.standard_normal((B, cdf, T))).astype(np.float32)).requires_grad_(True)
        lens = rng.randint(2, T + 1, size=B)
        lens[0] = T
        mask = torch.from_numpy(np.arange(T)[None, :] >= lens[:, None])
        att.applyMask(mask)
        wc, a = att(h, ctx)
        g = torch.from_numpy(rng.standard_normal(tuple(wc.shape)).astype(np.float32))
        wc.backward(g)
        for k, v in {"w": w, "h": h, "ctx": ctx, "g": g, "wc": wc, "attn": a, "dh": h.grad, "dctx": ctx.grad,
                     "dw": att.conv_context.weight.grad}.items():
            E["%s/%s" % (tag, k)] = summarize(v)
            E["%s/%s" % (tag, k)]["full"] = v.detach().numpy().reshape(-1).astype(np.float32)
        E["%s/mask" % tag] = summarize(mask.float())
        E["%s/mask" % tag]["full"] = mask.float().numpy().reshape(-1)
    # DAMSM losses
    cfg.TRAIN.SMOOTH.GAMMA1, cfg.TRAIN.SMOOTH.GAMMA2, cfg.TRAIN.SMOOTH.GAMMA3 = 4.0, 5.0, 10.0
    cfg.CUDA = False
    B, nef, T = 6, 32, 9
    feat = torch.from_numpy(rng.standard_normal((B, nef, 17, 17)).astype(np.float32)).requires_grad_(True)
    code = torch.from_numpy(rng.standard_normal((B, nef)).astype(np.float32)).requires_grad_(True)
    words = torch.from_numpy(np.tanh(rng.standard_normal((B, nef, T))).astype(np.float32))
    sent = torch.from_numpy(np.tanh(rng.standard_normal((B, nef))).astype(np.float32))
    lens = torch.from_numpy(np.array([9, 8, 8, 6, 5, 3]))
    labels = torch.arange(B)
    w0, w1, _ = L.words_loss(feat, words, labels, lens, np.arange(B), B)
    s0, s1 = L.sent_loss(code, sent, labels, np.arange(B), B)
    (w0 + w1 + s0 + s1).backward()
    for k, v in {"feat": feat, "code": code, "words": words, "sent": sent, "lens": lens.float(),
                 "w0": w0, "w1": w1, "s0": s0, "s1": s1, "dfeat": feat.grad, "dcode": code.grad}.items():
        E["damsm/%s" % k] = summarize(v)
        E["damsm/%s" % k]["full"] = v.detach().numpy().reshape(-1).astype(np.float32)
    save(name, E, {"what": "reference GlobalAttentionGeneral fwd/bwd, words_loss, sent_loss"})


def stn_cases(name):
    """The reference's stn() (model.py:17-21) on the four shapes it is used with, incl. an empty
    slot (all -1 bbox) and the 15x15 -> 16x16 resample of D_NET64 (quirk 7)."""
    E = {}
    rng = np.random.RandomState(11)
    bbox = np.array([[0.1, 0.2, 0.5, 0.4], [-1, -1, -1, -1], [0.55, 0.05, 0.4, 0.9], [0.0, 0.0, 0.998, 0.998]], np.float32)
    th = torch.from_numpy(synth.transformation_matrix(bbox))
    thi = torch.from_numpy(synth.transformation_matrix_inverse(bbox))
    B = bbox.shape[0]
    cases = {"scatter16": (thi, (B, 6, 16, 16), (B, 6, 16, 16)),
             "crop64to16": (th, (B, 3, 64, 64), (B, 3, 16, 16)),
             "scatter15to16": (thi, (B, 5, 15, 15), (B, 5, 16, 16)),
             "crop256to32": (th, (B, 3, 256, 256), (B, 3, 32, 32))}
    for tag, (theta, ishape, oshape) in cases.items():
        x = torch.from_numpy(rng.standard_normal(ishape).astype(np.float32)).requires_grad_(True)
        y = M.stn(x, theta, oshape)
        g = torch.from_numpy(rng.standard_normal(oshape).astype(np.float32))
        y.backward(g)
        for k, v in {"x": x, "theta": theta, "y": y, "g": g, "dx": x.grad}.items():
            E["%s/%s" % (tag, k)] = summarize(v)
            if k != "x" or x.numel() <= 65536:
                E["%s/%s" % (tag, k)]["full"] = v.detach().numpy().reshape(-1).astype(np.float32)
    E["bbox"] = summarize(torch.from_numpy(bbox))
    E["bbox"]["full"] = bbox.reshape(-1)
    # reference's own theta code (miscc/utils.py:16-49)
    E["theta_ref"] = summarize(U.compute_transformation_matrix(torch.from_numpy(bbox)))
    E["theta_inv_ref"] = summarize(U.compute_transformation_matrix_inverse(torch.from_numpy(bbox)))
    save(name, E, {"what": "reference stn fwd/bwd"})


TINY = dict(GF_DIM=8, DF_DIM=8, Z_DIM=20, R_NUM=2, EMBEDDING_DIM=32, T=6, B=4)

if __name__ == "__main__":
    torch.set_num_threads(8)
    import warnings
    warnings.filterwarnings("ignore")
    keys = attngan_step("attngan_tiny_step", TINY, seed=100)
    attngan_gd_only("attngan_tiny_gd", TINY, seed=100)
    attention_cases("attention_cases")
    stn_cases("stn_cases")
    # state_dict contract at the real config 5 dims (cfg/coco_train.yml)
    full = dict(GF_DIM=48, DF_DIM=96, Z_DIM=100, R_NUM=3, EMBEDDING_DIM=256, T=18, B=2)
    set_cfg(full)
    kfull = {"G_NET": {k: list(v.shape) for k, v in M.G_NET().state_dict().items()}}
    for i, cls in enumerate([M.D_NET64, M.D_NET128, M.D_NET256]):
        kfull["D_NET%d" % (64 << i)] = {k: list(v.shape) for k, v in cls().state_dict().items()}
    with open(os.path.join(HERE, "attngan_state_dict_keys.json"), "w") as f:
        json.dump({"tiny": keys, "config5": kfull}, f, indent=0)  # insertion order = parameter order
    print("done")
