"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product path.

CPU fp32 restatement (plain torch ops, functional style over a ``state_dict``) of the reference's
AttnGAN generator / discriminator / loss hot path.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this file; the product
package (``multiple-objects-gan_b200/mog_b200``) never does and fails loudly without its CUDA
library.

Parity pin: the reference has no tests or golden vectors of its own, so this
restatement is pinned against the *reference modules themselves*, executed in the build
container by ``tests/golden/make_golden.py`` (which imports ``/root/reference/code/coco/attngan``
unmodified); the resulting vectors are committed under ``tests/golden/`` and re-checked by
``tests/test_oracle_golden.py`` on every run.

Every function cites the reference lines it follows (paths relative to
``/root/reference/code/coco/attngan/``).  ``P`` is a dict ``name -> tensor`` using the
reference's ``state_dict`` key names; BatchNorm running statistics in ``P`` are updated in place
exactly like ``nn.BatchNorm*d`` in train mode would.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch
import torch.nn.functional as F

MAX_OBJECTS = 3  # model.py:14


@dataclass
class Cfg:
    """The cfg keys the hot path reads (miscc/config.py:23-64, cfg/coco_train.yml)."""
    GF_DIM: int = 48
    DF_DIM: int = 96
    Z_DIM: int = 100
    CONDITION_DIM: int = 100
    R_NUM: int = 3
    EMBEDDING_DIM: int = 256
    BRANCH_NUM: int = 3
    GAMMA1: float = 4.0
    GAMMA2: float = 5.0
    GAMMA3: float = 10.0
    LAMBDA: float = 50.0


# ------------------------------------------------------------------------------------------
# primitives
# ------------------------------------------------------------------------------------------
def stn(image, theta, size):
    """model.py:17-21 -- affine_grid + bilinear grid_sample, zero padding.  The container's torch
    defaults to align_corners=False; stated explicitly here."""
    grid = F.affine_grid(theta, list(size), align_corners=False)
    return F.grid_sample(image, grid, mode="bilinear", padding_mode="zeros", align_corners=False)


def glu(x):
    """model.py:24-32"""
    nc = x.size(1) // 2
    return x[:, :nc] * torch.sigmoid(x[:, nc:])


def batch_norm(x, P, prefix, train=True):
    """nn.BatchNorm1d/2d train mode: eps 1e-5, momentum 0.1, running stats updated in P."""
    nbt = P.get(prefix + ".num_batches_tracked")
    if train and nbt is not None:
        nbt += 1
    return F.batch_norm(x, P[prefix + ".running_mean"], P[prefix + ".running_var"],
                        P[prefix + ".weight"], P[prefix + ".bias"], training=train,
                        momentum=0.1, eps=1e-5)


def up_block(x, P, prefix):
    """model.py:48-55 -- nearest x2, conv3x3 (bias-free), BN, GLU.  Sequential indices 1, 2."""
    x = F.interpolate(x, scale_factor=2, mode="nearest")
    x = F.conv2d(x, P[prefix + ".1.weight"], None, 1, 1)
    x = batch_norm(x, P, prefix + ".2")
    return glu(x)


def res_block(x, P, prefix):
    """model.py:67-81 -- conv C->2C, BN, GLU, conv C->C, BN, += residual (no final activation)."""
    out = F.conv2d(x, P[prefix + ".block.0.weight"], None, 1, 1)
    out = glu(batch_norm(out, P, prefix + ".block.1"))
    out = F.conv2d(out, P[prefix + ".block.3.weight"], None, 1, 1)
    out = batch_norm(out, P, prefix + ".block.4")
    return out + x


def bbox_net(labels, theta_inv, P, prefix, c_dim):
    """model.py:105-116 -- label layout (sum over objects of the label vector painted into its
    box by the STN) followed by three stride-2 conv3x3 (model.py:90-103)."""
    B = labels.shape[0]
    layout = torch.zeros(B, c_dim, 16, 16)
    for idx in range(MAX_OBJECTS):
        cur = labels[:, idx].reshape(B, c_dim, 1, 1).repeat(1, 1, 16, 16)
        layout = layout + stn(cur, theta_inv[:, idx], cur.shape)
    x = F.leaky_relu(F.conv2d(layout, P[prefix + ".encode.0.weight"], None, 2, 1), 0.2)
    x = F.conv2d(x, P[prefix + ".encode.2.weight"], None, 2, 1)
    x = F.leaky_relu(batch_norm(x, P, prefix + ".encode.3"), 0.2)
    x = F.conv2d(x, P[prefix + ".encode.5.weight"], None, 2, 1)
    x = F.leaky_relu(batch_norm(x, P, prefix + ".encode.6"), 0.2)
    return x.reshape(B, -1)


def ca_net(sent_emb, P, cfg, eps=None):
    """model.py:317-345 -- Linear(+bias), GLU, split mu/logvar, c = mu + eps*exp(logvar/2).
    ``eps`` may be injected (parity tests); otherwise drawn like model.py:338."""
    x = glu(F.linear(sent_emb, P["ca_net.fc.weight"], P["ca_net.fc.bias"]))
    mu, logvar = x[:, :cfg.CONDITION_DIM], x[:, cfg.CONDITION_DIM:]
    std = (logvar * 0.5).exp()
    if eps is None:
        eps = torch.FloatTensor(std.size()).normal_()
    return eps * std + mu, mu, logvar


def init_stage_g(z_code, c_code, theta_inv, label_one_hot, P, cfg, prefix="h_net1"):
    """model.py:382-422.  Object pathway: per object idx the same label/local1/local2 modules
    are called with their own batch statistics (quirk 3), placed by the STN and summed."""
    B = z_code.shape[0]
    ngf = cfg.GF_DIM * 16
    ef = 100  # model.py:359
    local_labels = []
    h_locals = torch.zeros(B, ngf // 4, 16, 16)
    for idx in range(MAX_OBJECTS):
        lab = F.linear(torch.cat((c_code, label_one_hot[:, idx]), 1), P[prefix + ".label.0.weight"])
        lab = F.relu(batch_norm(lab, P, prefix + ".label.1"))
        local_labels.append(lab)
        h = lab.reshape(B, ef, 1, 1).repeat(1, 1, 4, 4)
        h = up_block(h, P, prefix + ".local1")
        h = up_block(h, P, prefix + ".local2")
        h_locals = h_locals + stn(h, theta_inv[:, idx], h.shape)
    local_labels = torch.stack(local_labels, 1)  # model.py:395 (copy-slices keeps the graph)
    bbox_code = bbox_net(local_labels, theta_inv, P, prefix + ".bbox_net", cfg.CONDITION_DIM)
    czc = torch.cat((c_code, z_code, bbox_code), 1)
    out = F.linear(czc, P[prefix + ".fc.0.weight"])
    out = glu(batch_norm(out, P, prefix + ".fc.1")).reshape(-1, ngf, 4, 4)
    out = up_block(out, P, prefix + ".upsample1")
    out = up_block(out, P, prefix + ".upsample2")
    out = torch.cat((out, h_locals), 1)
    out = up_block(out, P, prefix + ".upsample3")
    return up_block(out, P, prefix + ".upsample4")


def global_attention(h, context, mask, P, prefix):
    """GlobalAttention.py:82-123 including the mask-tiling quirk (``mask.repeat(queryL, 1)``
    against a batch-major (B*queryL, T) view, lines 104-108; applied on .data, no grad node)."""
    B, idf, ih, iw = h.shape
    queryL = ih * iw
    T = context.size(2)
    targetT = h.reshape(B, idf, queryL).transpose(1, 2).contiguous()
    sourceT = F.conv2d(context.unsqueeze(3), P[prefix + ".conv_context.weight"]).squeeze(3)
    attn = torch.bmm(targetT, sourceT).reshape(B * queryL, T)
    if mask is not None:
        with torch.no_grad():
            attn.masked_fill_(mask.repeat(queryL, 1), -float("inf"))
    attn = F.softmax(attn, dim=1).reshape(B, queryL, T).transpose(1, 2).contiguous()
    wc = torch.bmm(sourceT, attn).reshape(B, idf, ih, iw)
    return wc, attn.reshape(B, T, ih, iw)


def next_stage_g(h_code, word_embs, mask, P, cfg, prefix):
    """model.py:446-461 -- attention, concat, R_NUM ResBlocks, upBlock."""
    c_code, att = global_attention(h_code, word_embs, mask, P, prefix + ".att")
    out = torch.cat((h_code, c_code), 1)
    for i in range(cfg.R_NUM):
        out = res_block(out, P, "%s.residual.%d" % (prefix, i))
    return up_block(out, P, prefix + ".upsample"), att


def get_image_g(h, P, prefix):
    """model.py:464-475 -- conv3x3 -> 3 channels, tanh."""
    return torch.tanh(F.conv2d(h, P[prefix + ".img.0.weight"], None, 1, 1))


def g_net(P, cfg, z_code, sent_emb, word_embs, mask, theta_inv, label_one_hot, eps=None):
    """model.py:497-528 -> (fake_imgs[3], att_maps[2], mu, logvar)."""
    fake_imgs, att_maps = [], []
    c_code, mu, logvar = ca_net(sent_emb, P, cfg, eps)
    h1 = init_stage_g(z_code, c_code, theta_inv, label_one_hot, P, cfg)
    fake_imgs.append(get_image_g(h1, P, "img_net1"))
    h = h1
    for stage in range(2, cfg.BRANCH_NUM + 1):
        h, att = next_stage_g(h, word_embs, mask, P, cfg, "h_net%d" % stage)
        fake_imgs.append(get_image_g(h, P, "img_net%d" % stage))
        att_maps.append(att)
    return fake_imgs, att_maps, mu, logvar


# ------------------------------------------------------------------------------------------
# discriminators
# ------------------------------------------------------------------------------------------
def _down(x, P, conv, bn):
    x = F.conv2d(x, P[conv + ".weight"], None, 2, 1)
    return F.leaky_relu(batch_norm(x, P, bn), 0.2)


def _block3x3_leaky(x, P, prefix):
    """model.py:575-581"""
    x = F.conv2d(x, P[prefix + ".0.weight"], None, 1, 1)
    return F.leaky_relu(batch_norm(x, P, prefix + ".1"), 0.2)


def encode_image_by_16times(x, P, prefix):
    """model.py:595-613"""
    x = F.leaky_relu(F.conv2d(x, P[prefix + ".0.weight"], None, 2, 1), 0.2)
    x = _down(x, P, prefix + ".2", prefix + ".3")
    x = _down(x, P, prefix + ".5", prefix + ".6")
    return _down(x, P, prefix + ".8", prefix + ".9")


def d_net64(P, cfg, image, label, theta, theta_inv):
    """model.py:682-711 -- object pathway in D (crop by theta, concat label, 4x4/s1 conv to
    15x15, STN back to 16x16 by theta^-1, sum) + global pathway."""
    B = image.shape[0]
    ndf = cfg.DF_DIM
    h_locals = torch.zeros(B, ndf * 2, 16, 16)
    for idx in range(MAX_OBJECTS):
        lab = label[:, idx].reshape(B, 81, 1, 1).repeat(1, 1, 16, 16)
        h = stn(image, theta[:, idx], (B, image.shape[1], 16, 16))
        h = torch.cat((h, lab), 1)
        h = F.conv2d(h, P["local.0.weight"], None, 1, 1)
        h = F.leaky_relu(batch_norm(h, P, "local.1"), 0.2)
        h_locals = h_locals + stn(h, theta_inv[:, idx], (B, h.shape[1], 16, 16))
    h = F.leaky_relu(F.conv2d(image, P["conv1.weight"], None, 2, 1), 0.2)
    h = _down(h, P, "conv2", "bn2")
    h = torch.cat((h, h_locals), 1)
    h = _down(h, P, "conv3", "bn3")
    return _down(h, P, "conv4", "bn4")


def d_net128(P, cfg, x):
    """model.py:730-734"""
    x = encode_image_by_16times(x, P, "img_code_s16")
    x = _down(x, P, "img_code_s32.0", "img_code_s32.1")
    return _block3x3_leaky(x, P, "img_code_s32_1")


def d_net256(P, cfg, x):
    """model.py:754-760"""
    x = encode_image_by_16times(x, P, "img_code_s16")
    x = _down(x, P, "img_code_s32.0", "img_code_s32.1")
    x = _down(x, P, "img_code_s64.0", "img_code_s64.1")
    x = _block3x3_leaky(x, P, "img_code_s64_1")
    return _block3x3_leaky(x, P, "img_code_s64_2")


def d_get_logits(P, prefix, h_code, c_code=None):
    """model.py:629-642 -- (optional) condition concat + jointConv, then 4x4/s4 conv + Sigmoid."""
    if c_code is not None:
        c = c_code.reshape(c_code.shape[0], -1, 1, 1).repeat(1, 1, 4, 4)
        h_code = _block3x3_leaky(torch.cat((h_code, c), 1), P, prefix + ".jointConv")
    out = F.conv2d(h_code, P[prefix + ".outlogits.0.weight"], P[prefix + ".outlogits.0.bias"], 4)
    return torch.sigmoid(out).reshape(-1)


def d_features(which, P, cfg, img, label=None, theta=None, theta_inv=None):
    if which == 0:
        return d_net64(P, cfg, img, label, theta, theta_inv)
    return d_net128(P, cfg, img) if which == 1 else d_net256(P, cfg, img)


# ------------------------------------------------------------------------------------------
# losses
# ------------------------------------------------------------------------------------------
def discriminator_loss(which, P, cfg, real_img, fake_img, cond, real_labels, fake_labels,
                       label=None, theta=None, theta_inv=None):
    """miscc/losses.py:136-174 (UNCOND_DNET present, b_jcu=True)."""
    real_f = d_features(which, P, cfg, real_img, label, theta, theta_inv)
    fake_f = d_features(which, P, cfg, fake_img.detach(), label, theta, theta_inv)
    bce = F.binary_cross_entropy
    cond_real = bce(d_get_logits(P, "COND_DNET", real_f, cond), real_labels)
    cond_fake = bce(d_get_logits(P, "COND_DNET", fake_f, cond), fake_labels)
    B = real_f.size(0)
    cond_wrong = bce(d_get_logits(P, "COND_DNET", real_f[:B - 1], cond[1:B]), fake_labels[1:B])
    real = bce(d_get_logits(P, "UNCOND_DNET", real_f), real_labels)
    fake = bce(d_get_logits(P, "UNCOND_DNET", fake_f), fake_labels)
    return (real + cond_real) / 2.0 + (fake + cond_fake + cond_wrong) / 3.0


def generator_gan_loss(PDs, cfg, fake_imgs, sent_emb, real_labels, label, theta, theta_inv):
    """The adversarial part of miscc/losses.py:177-204: sum over D_i of uncond + cond BCE
    against the real labels (DAMSM terms handled by words_loss/sent_loss below)."""
    total = 0
    for i, PD in enumerate(PDs):
        f = d_features(i, PD, cfg, fake_imgs[i], label, theta, theta_inv)
        cond = F.binary_cross_entropy(d_get_logits(PD, "COND_DNET", f, sent_emb), real_labels)
        unc = F.binary_cross_entropy(d_get_logits(PD, "UNCOND_DNET", f), real_labels)
        total = total + unc + cond
    return total


def kl_loss(mu, logvar):
    """miscc/losses.py:230-234"""
    return torch.mean(1 + logvar - mu.pow(2) - logvar.exp()) * (-0.5)


def cosine_similarity(x1, x2, dim=1, eps=1e-8):
    """miscc/losses.py:11-17"""
    w12 = torch.sum(x1 * x2, dim)
    w1 = torch.norm(x1, 2, dim)
    w2 = torch.norm(x2, 2, dim)
    return (w12 / (w1 * w2).clamp(min=eps)).squeeze()


def func_attention(query, context, gamma1):
    """GlobalAttention.py:31-69 -- query B x ndf x queryL (words), context B x ndf x ih x iw."""
    B, queryL = query.size(0), query.size(2)
    ih, iw = context.size(2), context.size(3)
    sourceL = ih * iw
    context = context.reshape(B, -1, sourceL)
    contextT = context.transpose(1, 2).contiguous()
    attn = torch.bmm(contextT, query).reshape(B * sourceL, queryL)
    attn = F.softmax(attn, dim=1).reshape(B, sourceL, queryL)
    attn = attn.transpose(1, 2).contiguous().reshape(B * queryL, sourceL)
    attn = F.softmax(attn * gamma1, dim=1).reshape(B, queryL, sourceL)
    attnT = attn.transpose(1, 2).contiguous()
    return torch.bmm(context, attnT), attn.reshape(B, -1, ih, iw)


def _class_masks(class_ids, B):
    masks = np.zeros((B, B), bool)
    for i in range(B):
        m = (class_ids == class_ids[i])
        m[i] = False
        masks[i] = m
    return torch.from_numpy(masks)


def words_loss(img_features, words_emb, labels, cap_lens, class_ids, B, cfg):
    """miscc/losses.py:62-132"""
    sims = []
    lens = [int(v) for v in cap_lens]
    for i in range(B):
        n = lens[i]
        word = words_emb[i, :, :n].unsqueeze(0).repeat(B, 1, 1)
        wei, _ = func_attention(word, img_features, cfg.GAMMA1)
        w = word.transpose(1, 2).reshape(B * n, -1)
        c = wei.transpose(1, 2).reshape(B * n, -1)
        row = cosine_similarity(w, c).reshape(B, n)
        row = torch.log((row * cfg.GAMMA2).exp().sum(dim=1, keepdim=True))
        sims.append(row)
    sims = torch.cat(sims, 1) * cfg.GAMMA3
    if class_ids is not None:
        sims = sims.masked_fill(_class_masks(class_ids, B), -float("inf"))
    return F.cross_entropy(sims, labels), F.cross_entropy(sims.t(), labels)


def sent_loss(cnn_code, rnn_code, labels, class_ids, B, cfg, eps=1e-8):
    """miscc/losses.py:20-59"""
    cn = torch.norm(cnn_code, 2, dim=1, keepdim=True)
    rn = torch.norm(rnn_code, 2, dim=1, keepdim=True)
    scores = cnn_code @ rnn_code.t() / (cn @ rn.t()).clamp(min=eps) * cfg.GAMMA3
    if class_ids is not None:
        scores = scores.masked_fill(_class_masks(class_ids, B), -float("inf"))
    return F.cross_entropy(scores, labels), F.cross_entropy(scores.t(), labels)


# ------------------------------------------------------------------------------------------
# one G+D training step (forward + backward), trainer.py:294-340 without DAMSM / optimisers
# ------------------------------------------------------------------------------------------
def leafify(sd):
    """state_dict -> dict of leaf tensors (float params require grad, buffers cloned)."""
    P = {}
    for k, v in sd.items():
        v = v.detach().clone().cpu()
        if v.is_floating_point() and not k.endswith(("running_mean", "running_var")):
            v.requires_grad_(True)
        P[k] = v
    return P


def gd_step(PG, PDs, cfg, batch, eps=None, PE=None, image_encoder=None, after_d=None, noise=None):
    """G forward, three D losses + backward, G adversarial (+ DAMSM when the image-encoder weights ``PE`` are given,
    losses.py:205-224) + KL loss + backward.
    Returns losses, fake images and leaves gradients in ``.grad`` of PG / PDs[i].
    Order follows trainer.py:294-340.  D weight gradients produced by the G-step backward are
    *not* accumulated (they are discarded by the next zero_grad in the reference, trainer.py:304).
    ``image_encoder``: a callable img -> (region features, cnn_code) used instead of the Inception-v3 restatement;
    ``after_d(i)``: called once discriminator i's gradients are in place (trainer.py:317 ``optimizersD[i].step()``)."""
    B = batch["noise"].shape[0]
    noise = batch["noise"] if noise is None else noise
    real_labels, fake_labels = torch.ones(B), torch.zeros(B)
    fake_imgs, _, mu, logvar = g_net(PG, cfg, noise, batch["sent_emb"], batch["words_embs"],
                                     batch["mask"], batch["transf_matrices_inv"],
                                     batch["label_one_hot"], eps=eps if eps is not None else batch.get("eps"))
    errDs = []
    for i, PD in enumerate(PDs):
        kw = dict(label=batch["label_one_hot"], theta=batch["transf_matrices"],
                  theta_inv=batch["transf_matrices_inv"]) if i == 0 else {}
        errD = discriminator_loss(i, PD, cfg, batch["imgs"][i], fake_imgs[i], batch["sent_emb"],
                                  real_labels, fake_labels, **kw)
        params = [p for p in PD.values() if p.requires_grad]
        grads = torch.autograd.grad(errD, params)
        for p, g in zip(params, grads):
            p.grad = g
        errDs.append(errD.detach())
        if after_d is not None:
            after_d(i)
    errG = generator_gan_loss(PDs, cfg, fake_imgs, batch["sent_emb"], real_labels,
                              batch["label_one_hot"], batch["transf_matrices"],
                              batch["transf_matrices_inv"])
    if PE is not None or image_encoder is not None:
        match = torch.arange(B)
        if image_encoder is not None:
            region, code = image_encoder(fake_imgs[-1])
        else:
            from .encoder_oracle import cnn_encoder
            region, code = cnn_encoder(PE, fake_imgs[-1])
        w0, w1 = words_loss(region, batch["words_embs"], match, batch["cap_lens"], batch["class_ids"], B, cfg)[:2]
        s0, s1 = sent_loss(code, batch["sent_emb"], match, batch["class_ids"], B, cfg)
        errG = errG + (w0 + w1) * cfg.LAMBDA + (s0 + s1) * cfg.LAMBDA
    kl = kl_loss(mu, logvar)
    gparams = [p for p in PG.values() if p.requires_grad]
    grads = torch.autograd.grad(errG + kl, gparams, allow_unused=True)
    for p, g in zip(gparams, grads):
        p.grad = g
    return {"errD": errDs, "errG": errG.detach(), "kl": kl.detach(),
            "fake_imgs": [f.detach() for f in fake_imgs], "mu": mu.detach(), "logvar": logvar.detach()}


def make_train_state(PG, PDs, lr=2e-4):
    """trainer.py:139-160,251 -- Adam(lr, betas=(0.5, 0.999)) per network and the EMA copy of the G parameters."""
    gp = [p for p in PG.values() if p.requires_grad]
    return {"optG": torch.optim.Adam(gp, lr=lr, betas=(0.5, 0.999)),
            "optDs": [torch.optim.Adam([p for p in PD.values() if p.requires_grad], lr=lr, betas=(0.5, 0.999)) for PD in PDs],
            "ema": [p.detach().clone() for p in gp], "gparams": gp}


def train_step(PG, PDs, state, cfg, batch, eps=None, noise=None, PE=None, image_encoder=None):
    """One iteration of trainer.py:294-342 including the optimiser steps (each D right after its backward, so the G
    step sees the UPDATED discriminators) and the EMA of the generator (trainer.py:341-342)."""
    def after_d(i):
        state["optDs"][i].step()
    out = gd_step(PG, PDs, cfg, batch, eps=eps, PE=PE, image_encoder=image_encoder, after_d=after_d, noise=noise)
    state["optG"].step()
    with torch.no_grad():
        for p, a in zip(state["gparams"], state["ema"]):
            a.mul_(0.999).add_(p.detach(), alpha=0.001)
    return out
