"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product path.

CPU fp32 restatement (plain torch ops over a ``state_dict``) of the reference's COCO StackGAN
program: ``code/coco/stackgan/model.py`` (STAGE1_G/D, STAGE2_G/D, CA_NET, BBOX_NET, D_GET_LOGITS,
ResBlock) and the losses of ``code/coco/stackgan/miscc/utils.py:68-125``.  Pinned against vectors
produced by executing the unmodified reference (``tests/golden/make_golden_stackgan.py`` ->
``tests/golden/stackgan_s{1,2}.npz``) by ``tests/test_stackgan.py``.  Only tests / smoke / bench CPU
legs may import it.  Line numbers below are those of ``stackgan/model.py`` unless stated.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .attngan_oracle import batch_norm, stn

N_LABELS = 81


def up_block(x, P, prefix):
    """:16-22 -- nearest x2, conv3x3, BN, ReLU."""
    x = F.interpolate(x, scale_factor=2, mode="nearest")
    x = F.conv2d(x, P[prefix + ".1.weight"], None, 1, 1)
    return F.relu(batch_norm(x, P, prefix + ".2"))


def res_block(x, P, prefix):
    """:25-41 -- conv BN ReLU conv BN, += residual, ReLU."""
    h = F.relu(batch_norm(F.conv2d(x, P[prefix + ".block.0.weight"], None, 1, 1), P, prefix + ".block.1"))
    h = batch_norm(F.conv2d(h, P[prefix + ".block.3.weight"], None, 1, 1), P, prefix + ".block.4")
    return F.relu(h + x)


def ca_net(P, prefix, text_embedding, eps):
    """:44-73 -- eps is the reference's ``FloatTensor(std.size()).normal_()`` draw, injected."""
    x = F.relu(F.linear(text_embedding, P[prefix + ".fc.weight"], P[prefix + ".fc.bias"]))
    c = x.shape[1] // 2
    mu, logvar = x[:, :c], x[:, c:]
    return eps * torch.exp(0.5 * logvar) + mu, mu, logvar


def label_layout(labels, theta_inv, n_obj):
    """:141-147 / :395-405 -- sum over objects of the spatially replicated label placed by theta^-1."""
    B, C = labels.shape[0], labels.shape[2]
    layout = torch.zeros(B, C, 16, 16)
    for idx in range(n_obj):
        cur = labels[:, idx].reshape(B, C, 1, 1).repeat(1, 1, 16, 16)
        layout = layout + stn(cur, theta_inv[:, idx], cur.shape)
    return layout


def bbox_net(labels, theta_inv, P, prefix, n_obj):
    """:114-149"""
    x = label_layout(labels, theta_inv, n_obj)
    x = F.leaky_relu(F.conv2d(x, P[prefix + ".encode.0.weight"], None, 2, 1), 0.2)
    x = F.leaky_relu(batch_norm(F.conv2d(x, P[prefix + ".encode.2.weight"], None, 2, 1), P, prefix + ".encode.3"), 0.2)
    x = F.leaky_relu(batch_norm(F.conv2d(x, P[prefix + ".encode.5.weight"], None, 2, 1), P, prefix + ".encode.6"), 0.2)
    return x.reshape(labels.shape[0], -1)


def object_label(P, prefix, c_code, one_hot):
    """``self.label(cat(c_code, label_one_hot[:, idx]))`` (:214) -- Linear, BN1d (own batch statistics), ReLU."""
    h = F.linear(torch.cat((c_code, one_hot.float()), 1), P[prefix + ".0.weight"])
    return F.relu(batch_norm(h, P, prefix + ".1"))


def stage1_g(P, text_embedding, noise, theta_inv, label_one_hot, eps, gf_dim, n_obj=3, use_bbox=True, pre=""):
    """:194-245.  gf_dim = cfg.GAN.GF_DIM * 8.  Returns (fake_img, mu, logvar, local_labels)."""
    B = noise.shape[0]
    c_code, mu, logvar = ca_net(P, pre + "ca_net", text_embedding, eps)
    h_locals = torch.zeros(B, gf_dim // 4, 16, 16)
    labs = []
    for idx in range(n_obj):
        lab = object_label(P, pre + "label", c_code, label_one_hot[:, idx])
        labs.append(lab)
        h = lab.reshape(B, -1, 1, 1).repeat(1, 1, 4, 4)
        h = up_block(h, P, pre + "local1")
        h = up_block(h, P, pre + "local2")
        h_locals = h_locals + stn(h, theta_inv[:, idx], h.shape)
    local_labels = torch.stack(labs, 1)
    if use_bbox:
        z_c = torch.cat((noise, c_code, bbox_net(local_labels, theta_inv, P, pre + "bbox_net", n_obj)), 1)
    else:
        z_c = torch.cat((noise, c_code), 1)
    h = F.relu(batch_norm(F.linear(z_c, P[pre + "fc.0.weight"]), P, pre + "fc.1")).reshape(-1, gf_dim, 4, 4)
    h = up_block(h, P, pre + "upsample1")
    h = up_block(h, P, pre + "upsample2")
    h = torch.cat((h, h_locals), 1)
    h = up_block(h, P, pre + "upsample3")
    h = up_block(h, P, pre + "upsample4")
    return torch.tanh(F.conv2d(h, P[pre + "img.0.weight"], None, 1, 1)), mu, logvar, local_labels


def _d_locals(P, image, label, theta, theta_inv, n_obj, size, n_local):
    """:271-286 / :487-501 -- crop, concat one-hot label planes, `local` convs (4x4 s1 p1), place back."""
    B = image.shape[0]
    canvas = None
    for idx in range(n_obj):
        lab = label[:, idx].float().reshape(B, N_LABELS, 1, 1).repeat(1, 1, size, size)
        h = stn(image, theta[:, idx], (B, image.shape[1], size, size))
        h = torch.cat((h, lab), 1)
        for j in range(n_local):
            h = F.conv2d(h, P["local.%d.weight" % (3 * j)], None, 1, 1)
            h = F.leaky_relu(batch_norm(h, P, "local.%d" % (3 * j + 1)), 0.2)
        h = stn(h, theta_inv[:, idx], (B, h.shape[1], size, size))
        canvas = h if canvas is None else canvas + h
    return canvas


def stage1_d(P, image, label, theta, theta_inv, n_obj=3):
    """:248-309"""
    h_locals = _d_locals(P, image, label, theta, theta_inv, n_obj, 16, 1)
    h = F.leaky_relu(F.conv2d(image, P["conv1.weight"], None, 2, 1), 0.2)
    h = F.leaky_relu(batch_norm(F.conv2d(h, P["conv2.weight"], None, 2, 1), P, "bn2"), 0.2)
    h = torch.cat((h, h_locals), 1)
    h = F.leaky_relu(batch_norm(F.conv2d(h, P["conv3.weight"], None, 2, 1), P, "bn3"), 0.2)
    return F.leaky_relu(batch_norm(F.conv2d(h, P["conv4.weight"], None, 2, 1), P, "bn4"), 0.2)


def stage2_g(P, text_embedding, noise, theta_inv, theta_s2, theta_inv_s2, label_one_hot, eps1, eps2, gf_dim, r_num,
             n_obj=3, use_bbox=True):
    """:377-444.  gf_dim = cfg.GAN.GF_DIM (192 for the hard-coded 768 of :340 to hold).
    Returns (stage1_img, fake_img, mu, logvar, local_labels)."""
    B = noise.shape[0]
    with torch.no_grad():   # frozen stage-I generator, train-mode BatchNorm (:379)
        s1_img, _, _, _ = stage1_g(P, text_embedding, noise, theta_inv, label_one_hot, eps1, gf_dim * 8, n_obj, use_bbox,
                                   pre="STAGE1_G.")
    x = F.relu(F.conv2d(s1_img, P["encoder.0.weight"], None, 1, 1))
    x = F.relu(batch_norm(F.conv2d(x, P["encoder.2.weight"], None, 2, 1), P, "encoder.3"))
    enc = F.relu(batch_norm(F.conv2d(x, P["encoder.5.weight"], None, 2, 1), P, "encoder.6"))
    c_code, mu, logvar = ca_net(P, "ca_net", text_embedding, eps2)
    c_rep = c_code.reshape(B, -1, 1, 1).repeat(1, 1, 16, 16)
    labs = [object_label(P, "label", c_code, label_one_hot[:, idx]) for idx in range(n_obj)]
    local_labels = torch.stack(labs, 1)
    if use_bbox:
        i_c = torch.cat((enc, c_rep, label_layout(local_labels, theta_inv, n_obj)), 1)
    else:
        i_c = torch.cat((enc, c_rep), 1)
    h = F.relu(batch_norm(F.conv2d(i_c, P["hr_joint.0.weight"], None, 1, 1), P, "hr_joint.1"))
    for i in range(r_num):
        h = res_block(h, P, "residual.%d" % i)
    h_locals = torch.zeros(B, gf_dim, 64, 64)
    for idx in range(n_obj):
        lab = local_labels[:, idx].reshape(B, -1, 1, 1).repeat(1, 1, 16, 16)
        patch = stn(h, theta_s2[:, idx], (B, h.shape[1], 16, 16))
        g = up_block(torch.cat((patch, lab), 1), P, "local1")
        g = up_block(g, P, "local2")
        h_locals = h_locals + stn(g, theta_inv_s2[:, idx], h_locals.shape)
    h = up_block(h, P, "upsample1")
    h = up_block(h, P, "upsample2")
    h = torch.cat((h, h_locals), 1)
    h = up_block(h, P, "upsample3")
    h = up_block(h, P, "upsample4")
    return s1_img, torch.tanh(F.conv2d(h, P["img.0.weight"], None, 1, 1)), mu, logvar, local_labels


def stage2_d(P, image, label, theta, theta_inv, n_obj=3):
    """:447-537"""
    h_locals = _d_locals(P, image, label, theta, theta_inv, n_obj, 32, 2)
    h = F.leaky_relu(F.conv2d(image, P["conv1.weight"], None, 2, 1), 0.2)
    h = F.leaky_relu(batch_norm(F.conv2d(h, P["conv2.weight"], None, 2, 1), P, "bn2"), 0.2)
    h = F.leaky_relu(batch_norm(F.conv2d(h, P["conv3.weight"], None, 2, 1), P, "bn3"), 0.2)
    h = torch.cat((h, h_locals), 1)
    for i, (s, p) in zip(range(4, 9), ((2, 1), (2, 1), (2, 1), (1, 1), (1, 1))):
        h = F.leaky_relu(batch_norm(F.conv2d(h, P["conv%d.weight" % i], None, s, p), P, "bn%d" % i), 0.2)
    return h


def cond_logits(P, h_code, cond):
    """D_GET_LOGITS bcondition=True (:76-104); no Sigmoid."""
    c = cond.reshape(cond.shape[0], -1, 1, 1).repeat(1, 1, 4, 4)
    h = F.conv2d(torch.cat((h_code, c), 1), P["get_cond_logits.outlogits.0.weight"], None, 1, 1)
    h = F.leaky_relu(batch_norm(h, P, "get_cond_logits.outlogits.1"), 0.2)
    return F.conv2d(h, P["get_cond_logits.outlogits.3.weight"], P["get_cond_logits.outlogits.3.bias"], 4).reshape(-1)


def uncond_logits(P, h_code):
    return F.conv2d(h_code, P["get_uncond_logits.outlogits.0.weight"], P["get_uncond_logits.outlogits.0.bias"], 4).reshape(-1)


def kl_loss(mu, logvar):
    """miscc/utils.py:68-71"""
    return -0.5 * torch.mean(1 + logvar - mu.pow(2) - logvar.exp())


def discriminator_loss(P, d_fn, has_uncond, real, fake, label, theta, theta_inv, cond):
    """miscc/utils.py:74-109"""
    B = real.shape[0]
    ones, zeros = torch.ones(B), torch.zeros(B)
    cond = cond.detach()
    rf = d_fn(P, real, label, theta, theta_inv)
    ff = d_fn(P, fake.detach(), label, theta, theta_inv)
    bce = F.binary_cross_entropy_with_logits
    e_real = bce(cond_logits(P, rf, cond), ones)
    e_wrong = bce(cond_logits(P, rf[:B - 1], cond[1:]), zeros[1:])
    e_fake = bce(cond_logits(P, ff, cond), zeros)
    if has_uncond:
        u_real = bce(uncond_logits(P, rf), ones)
        u_fake = bce(uncond_logits(P, ff), zeros)
        return (e_real + u_real) / 2. + (e_fake + e_wrong + u_fake) / 3.
    return e_real + (e_fake + e_wrong) * 0.5


def generator_loss(P, d_fn, has_uncond, fake, label, theta, theta_inv, cond):
    """miscc/utils.py:112-125"""
    ones = torch.ones(fake.shape[0])
    ff = d_fn(P, fake, label, theta, theta_inv)
    e = F.binary_cross_entropy_with_logits(cond_logits(P, ff, cond.detach()), ones)
    if has_uncond:
        e = e + F.binary_cross_entropy_with_logits(uncond_logits(P, ff), ones)
    return e
