"""GAN losses of the AttnGAN program -- libmog edition of ``code/coco/attngan/miscc/losses.py``
(same function names and argument lists).

The discriminator heads end in nn.Sigmoid followed by nn.BCELoss in the reference
(model.py:627, losses.py:156-171); here both are one fused kernel on the pre-sigmoid logits
(``mog_sigmoid_bce_*``, identical arithmetic incl. the log clamp at -100).
``nn.parallel.data_parallel(netD, inputs, gpus)`` becomes a direct call: this build is one
process per GPU with NCCL gradient all-reduce (see ``mog_b200/parallel.py``), so ``gpus`` is
accepted and ignored.
"""
from __future__ import annotations

import torch

from ... import ops
from .config import cfg


def _features(netD, imgs, local_labels, transf_matrices, transf_matrices_inv):
    if local_labels is not None:
        return netD(imgs, local_labels, transf_matrices, transf_matrices_inv)
    return netD(imgs)


def discriminator_loss(netD, real_imgs, fake_imgs, conditions, real_labels, fake_labels, gpus=None,
                       local_labels=None, transf_matrices=None, transf_matrices_inv=None):
    """losses.py:136-174"""
    real_features = _features(netD, real_imgs, local_labels, transf_matrices, transf_matrices_inv)
    fake_features = _features(netD, fake_imgs.detach(), local_labels, transf_matrices, transf_matrices_inv)
    bce = ops.sigmoid_bce
    cond_real_errD = bce(netD.COND_DNET.logits(real_features, conditions), real_labels)
    cond_fake_errD = bce(netD.COND_DNET.logits(fake_features, conditions), fake_labels)
    batch_size = real_features.size(0)
    cond_wrong_errD = bce(netD.COND_DNET.logits(real_features[:(batch_size - 1)], conditions[1:batch_size]),
                          fake_labels[1:batch_size])
    if netD.UNCOND_DNET is not None:
        real_errD = bce(netD.UNCOND_DNET.logits(real_features), real_labels)
        fake_errD = bce(netD.UNCOND_DNET.logits(fake_features), fake_labels)
        errD = ((real_errD + cond_real_errD) / 2. + (fake_errD + cond_fake_errD + cond_wrong_errD) / 3.)
    else:
        errD = cond_real_errD + (cond_fake_errD + cond_wrong_errD) / 2.
    return errD


def generator_loss(netsD, image_encoder, fake_imgs, real_labels, words_embs, sent_emb, match_labels,
                   cap_lens, class_ids, gpus=None, local_labels=None, transf_matrices=None,
                   transf_matrices_inv=None):
    """losses.py:177-226.  ``logs`` is returned as a list of (name, 0-d tensor) pairs instead of a
    formatted string so that no ``.item()`` host sync happens inside the step (the reference forces
    4-5 syncs per step, losses.py:204,225); ``format_logs`` renders the reference's string.
    With ``image_encoder is None`` the DAMSM terms (losses.py:205-224) are skipped (G+D-only step)."""
    numDs = len(netsD)
    logs = []
    errG_total = 0
    for i in range(numDs):
        if i == 0:
            features = netsD[i](fake_imgs[i], local_labels, transf_matrices, transf_matrices_inv)
        else:
            features = netsD[i](fake_imgs[i])
        cond_errG = ops.sigmoid_bce(netsD[i].COND_DNET.logits(features, sent_emb), real_labels)
        if netsD[i].UNCOND_DNET is not None:
            errG = ops.sigmoid_bce(netsD[i].UNCOND_DNET.logits(features), real_labels)
            g_loss = errG + cond_errG
        else:
            g_loss = cond_errG
        errG_total = errG_total + g_loss
        logs.append(('g_loss%d' % i, g_loss.detach()))
        if i == (numDs - 1) and image_encoder is not None:
            region_features, cnn_code = image_encoder(fake_imgs[i])
            batch_size = real_labels.size(0)
            w_loss0, w_loss1, _ = words_loss(region_features, words_embs, match_labels, cap_lens, class_ids, batch_size)
            w_loss = (w_loss0 + w_loss1) * cfg.TRAIN.SMOOTH.LAMBDA
            s_loss0, s_loss1 = sent_loss(cnn_code, sent_emb, match_labels, class_ids, batch_size)
            s_loss = (s_loss0 + s_loss1) * cfg.TRAIN.SMOOTH.LAMBDA
            errG_total = errG_total + w_loss + s_loss
            logs.append(('w_loss', w_loss.detach()))
            logs.append(('s_loss', s_loss.detach()))
    return errG_total, logs


def format_logs(logs):
    return ''.join('%s: %.2f ' % (k, float(v)) for k, v in logs)


def KL_loss(mu, logvar):
    """losses.py:230-234 (B x 100 elementwise: plain torch, not a hot-path kernel)."""
    KLD_element = mu.pow(2).add(logvar.exp()).mul(-1).add(1).add(logvar)
    return torch.mean(KLD_element).mul(-0.5)


def words_loss(img_features, words_emb, labels, cap_lens, class_ids, batch_size):
    raise NotImplementedError("DAMSM words_loss: fused all-pairs kernel is the next scope row (SURVEY 8 a18)")


def sent_loss(cnn_code, rnn_code, labels, class_ids, batch_size, eps=1e-8):
    raise NotImplementedError("DAMSM sent_loss: next scope row (SURVEY 8 a19)")
