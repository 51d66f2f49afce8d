"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product path.

CPU fp32 restatement (plain torch ops over a ``state_dict``) of the reference's single-stage
programs: Multi-MNIST (``code/multi-mnist/model.py``, ``miscc/utils.py:71-123``) and CLEVR
(``code/clevr/model.py``, ``miscc/utils.py:93-142``).  Pinned against vectors produced by executing
the unmodified reference (``tests/golden/make_golden_stage1.py`` -> ``tests/golden/stage1_*.npz``).
Only tests / smoke / bench CPU legs may import it.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.nn.functional as F

from .attngan_oracle import batch_norm, stn


@dataclass
class Flavor:
    name: str
    n_label: int
    img_ch: int
    n_objects: int
    embed_label: bool      # clevr/model.py:164: label -> Linear+BN1d+ReLU
    clamp_cond: bool       # clevr/miscc/utils.py:99


MNIST = Flavor("multi-mnist", 10, 1, 3, False, False)
CLEVR = Flavor("clevr", 13, 3, 4, True, True)


def up_block(x, P, prefix):
    """multi-mnist/model.py:16-22 -- nearest x2, conv3x3, BN, ReLU."""
    x = F.interpolate(x, scale_factor=2, mode="nearest")
    x = F.conv2d(x, P[prefix + ".1.weight"], None, 1, 1)
    return F.relu(batch_norm(x, P, prefix + ".2"))


def bbox_net(labels, theta_inv, P, prefix, n_obj):
    """multi-mnist/model.py:98-111"""
    B, C = labels.shape[0], labels.shape[2]
    layout = torch.zeros(B, C, 16, 16)
    for idx in range(n_obj):
        cur = labels[:, idx].reshape(B, C, 1, 1).repeat(1, 1, 16, 16)
        layout = layout + stn(cur, theta_inv[:, idx], cur.shape)
    x = F.leaky_relu(F.conv2d(layout, P[prefix + ".encode.0.weight"], None, 2, 1), 0.2)
    x = F.leaky_relu(batch_norm(F.conv2d(x, P[prefix + ".encode.2.weight"], None, 2, 1), P, prefix + ".encode.3"), 0.2)
    x = F.leaky_relu(batch_norm(F.conv2d(x, P[prefix + ".encode.5.weight"], None, 2, 1), P, prefix + ".encode.6"), 0.2)
    return x.reshape(B, -1)


def stage1_g(P, fl: Flavor, noise, theta_inv, label_one_hot, gf_dim):
    """multi-mnist/model.py:158-192 / clevr/model.py:158-194.  gf_dim = cfg.GAN.GF_DIM * 8."""
    B = noise.shape[0]
    h_locals = torch.zeros(B, gf_dim // 4, 16, 16)
    labs = []
    for idx in range(fl.n_objects):
        lab = label_one_hot[:, idx].float()
        if fl.embed_label:
            lab = F.relu(batch_norm(F.linear(lab, P["label.0.weight"]), P, "label.1"))
        labs.append(lab)
        h = lab.reshape(B, -1, 1, 1).repeat(1, 1, 4, 4)
        h = up_block(h, P, "local1")
        h = up_block(h, P, "local2")
        h_locals = h_locals + stn(h, theta_inv[:, idx], h.shape)
    bbox_code = bbox_net(torch.stack(labs, 1), theta_inv, P, "bbox_net", fl.n_objects)
    h = F.linear(torch.cat((noise, bbox_code), 1), P["fc.0.weight"])
    h = F.relu(batch_norm(h, P, "fc.1")).reshape(-1, gf_dim, 4, 4)
    h = up_block(h, P, "upsample1")
    h = up_block(h, P, "upsample2")
    h = torch.cat((h, h_locals), 1)
    h = up_block(h, P, "upsample3")
    h = up_block(h, P, "upsample4")
    return torch.tanh(F.conv2d(h, P["img.0.weight"], None, 1, 1))


def stage1_d(P, fl: Flavor, image, label, theta, theta_inv, df_dim):
    """multi-mnist/model.py:224-257"""
    B = image.shape[0]
    h_locals = torch.zeros(B, df_dim * 2, 16, 16)
    for idx in range(fl.n_objects):
        lab = label[:, idx].float().reshape(B, fl.n_label, 1, 1).repeat(1, 1, 16, 16)
        h = stn(image, theta[:, idx], (B, image.shape[1], 16, 16))
        h = F.conv2d(torch.cat((h, lab), 1), P["local.0.weight"], None, 1, 1)
        h = F.leaky_relu(batch_norm(h, P, "local.1"), 0.2)
        h_locals = h_locals + stn(h, theta_inv[:, idx], (B, h.shape[1], 16, 16))
    h = F.leaky_relu(F.conv2d(image, P["conv1.weight"], None, 2, 1), 0.2)
    h = F.leaky_relu(batch_norm(F.conv2d(h, P["conv2.weight"], None, 2, 1), P, "bn2"), 0.2)
    h = torch.cat((h, h_locals), 1)
    h = F.leaky_relu(batch_norm(F.conv2d(h, P["conv3.weight"], None, 2, 1), P, "bn3"), 0.2)
    return F.leaky_relu(batch_norm(F.conv2d(h, P["conv4.weight"], None, 2, 1), P, "bn4"), 0.2)


def cond_logits(P, h_code, cond):
    """D_GET_LOGITS (bcondition=True), multi-mnist/model.py:62-71"""
    c = cond.reshape(cond.shape[0], -1, 1, 1).repeat(1, 1, 4, 4)
    h = F.conv2d(torch.cat((h_code, c), 1), P["get_cond_logits.outlogits.0.weight"], None, 1, 1)
    h = F.leaky_relu(batch_norm(h, P, "get_cond_logits.outlogits.1"), 0.2)
    return F.conv2d(h, P["get_cond_logits.outlogits.3.weight"], P["get_cond_logits.outlogits.3.bias"], 4).reshape(-1)


def label_cond(fl, local_label):
    cond = sum(local_label[:, i, :].float() for i in range(fl.n_objects))
    return cond.clamp_min(0) if fl.clamp_cond else cond


def discriminator_loss(P, fl, real, fake, local_label, theta, theta_inv, df_dim):
    """multi-mnist/miscc/utils.py:71-109 (get_uncond_logits is None)."""
    B = real.shape[0]
    ones, zeros = torch.ones(B), torch.zeros(B)
    cond = label_cond(fl, local_label)
    rf = stage1_d(P, fl, real, local_label, theta, theta_inv, df_dim)
    ff = stage1_d(P, fl, fake.detach(), local_label, theta, theta_inv, df_dim)
    bce = F.binary_cross_entropy_with_logits
    e_real = bce(cond_logits(P, rf, cond), ones)
    e_wrong = bce(cond_logits(P, rf[:B - 1], cond[1:]), zeros[1:])
    e_fake = bce(cond_logits(P, ff, cond), zeros)
    return e_real + (e_fake + e_wrong) * 0.5


def generator_loss(P, fl, fake, local_label, theta, theta_inv, df_dim):
    """multi-mnist/miscc/utils.py:112-123"""
    cond = label_cond(fl, local_label)
    ff = stage1_d(P, fl, fake, local_label, theta, theta_inv, df_dim)
    return F.binary_cross_entropy_with_logits(cond_logits(P, ff, cond), torch.ones(fake.shape[0]))
