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


PAIR_PASS = True   # discriminator_loss: real + fake batch in one two-segment pass


def discriminator_loss(netD, real_imgs, fake_imgs, conditions, real_labels, fake_labels, gpus=None,
                       local_labels=None, transf_matrices=None, transf_matrices_inv=None):
    """losses.py:136-174.  The real and the fake batch (same size) go through the discriminator and its conditional
    head in ONE pass as two BatchNorm segments -- the statistics, running-stat updates and results of the reference's
    separate calls, with every weight streamed once (``PAIR_PASS = False`` restores the call-by-call form)."""
    bce = ops.sigmoid_bce
    batch_size = real_imgs.size(0)
    if PAIR_PASS and hasattr(netD, "forward_pair") and fake_imgs.shape == real_imgs.shape:
        if local_labels is not None:
            both = netD.forward_pair(real_imgs, fake_imgs.detach(), local_labels, transf_matrices, transf_matrices_inv)
        else:
            both = netD.forward_pair(real_imgs, fake_imgs.detach())
        real_features, fake_features = both[:batch_size], both[batch_size:]
        cond = netD.COND_DNET.logits(both, torch.cat((conditions, conditions), 0), segments=2)
        cond_real_errD = bce(cond[:batch_size], real_labels)
        cond_fake_errD = bce(cond[batch_size:], fake_labels)
        uncond = netD.UNCOND_DNET.logits(both) if netD.UNCOND_DNET is not None else None
    else:
        real_features = _features(netD, real_imgs, local_labels, transf_matrices, transf_matrices_inv)
        fake_features = _features(netD, fake_imgs.detach(), local_labels, transf_matrices, transf_matrices_inv)
        cond_real_errD = bce(netD.COND_DNET.logits(real_features, conditions), real_labels)
        cond_fake_errD = bce(netD.COND_DNET.logits(fake_features, conditions), fake_labels)
        uncond = None
    cond_wrong_errD = bce(netD.COND_DNET.logits(real_features[:(batch_size - 1)], conditions[1:batch_size]),
                          fake_labels[1:batch_size])
    if netD.UNCOND_DNET is not None:
        if uncond is not None:
            real_errD, fake_errD = bce(uncond[:batch_size], real_labels), bce(uncond[batch_size:], fake_labels)
        else:
            real_errD = bce(netD.UNCOND_DNET.logits(real_features), real_labels)
            fake_errD = bce(netD.UNCOND_DNET.logits(fake_features), fake_labels)
        errD = ((real_errD + cond_real_errD) / 2. + (fake_errD + cond_fake_errD + cond_wrong_errD) / 3.)
    else:
        errD = cond_real_errD + (cond_fake_errD + cond_wrong_errD) / 2.
    return errD


def damsm_terms(image_encoder, fake_img, words_embs, sent_emb, match_labels, cap_lens, class_ids, batch_size):
    """The DAMSM part of the generator objective (losses.py:205-224): (w_loss, s_loss), each already times LAMBDA.  It depends
    on the generated image and the frozen encoders only -- not on the discriminators -- so a trainer may evaluate it on a side
    stream while the discriminator steps run and hand it to :func:`generator_loss` (``damsm=``)."""
    region_features, cnn_code = image_encoder(fake_img)
    w_loss0, w_loss1, _ = words_loss(region_features, words_embs, match_labels, cap_lens, class_ids, batch_size)
    s_loss0, s_loss1 = sent_loss(cnn_code, sent_emb, match_labels, class_ids, batch_size)
    return (w_loss0 + w_loss1) * cfg.TRAIN.SMOOTH.LAMBDA, (s_loss0 + s_loss1) * cfg.TRAIN.SMOOTH.LAMBDA


def generator_loss(netsD, image_encoder, fake_imgs, real_labels, words_embs, sent_emb, match_labels,
                   cap_lens, class_ids, gpus=None, local_labels=None, transf_matrices=None,
                   transf_matrices_inv=None, streams=None, damsm=None):
    """losses.py:177-226.  ``logs`` is returned as a list of (name, 0-d tensor) pairs instead of a
    formatted string so that no ``.item()`` host sync happens inside the step (the reference forces
    4-5 syncs per step, losses.py:204,225); ``format_logs`` renders the reference's string.
    With ``image_encoder is None`` the DAMSM terms (losses.py:205-224) are skipped (G+D-only step)."""
    numDs = len(netsD)
    batch_size = real_labels.size(0)

    def d_branch(i):
        if i == 0:
            features = netsD[i](fake_imgs[i], local_labels, transf_matrices, transf_matrices_inv)
        else:
            features = netsD[i](fake_imgs[i])
        cond_errG = ops.sigmoid_bce(netsD[i].COND_DNET.logits(features, sent_emb), real_labels)
        if netsD[i].UNCOND_DNET is not None:
            return ops.sigmoid_bce(netsD[i].UNCOND_DNET.logits(features), real_labels) + cond_errG
        return cond_errG

    def damsm_branch():
        return damsm_terms(image_encoder, fake_imgs[numDs - 1], words_embs, sent_emb, match_labels, cap_lens, class_ids, batch_size)

    # The branches (one per discriminator, one through the image encoder) are independent until the final sum.  With
    # ``streams`` they are enqueued on separate CUDA streams (forked from / joined to the current one): the many small kernels
    # of the encoder and of the low-resolution discriminators fill the SMs / the HBM bandwidth the big tensor-core kernels
    # of D_NET256 leave idle, in the forward and -- autograd runs every node on its forward's stream -- in the backward.
    # Same kernels, same summation order: the result is bit-identical to the single-stream order.
    # ``damsm``: the terms already evaluated by the caller (on streams[numDs], joined below)
    g_losses = [None] * numDs
    if streams is None:
        for i in range(numDs):
            g_losses[i] = d_branch(i)
        if image_encoder is not None and damsm is None:
            damsm = damsm_branch()
    else:
        cur = torch.cuda.current_stream()
        for s in streams[:numDs]:
            s.wait_stream(cur)
        if image_encoder is not None and damsm is None:       # the longest chain of small kernels first
            streams[numDs].wait_stream(cur)
            with torch.cuda.stream(streams[numDs]):
                damsm = damsm_branch()
        for i in reversed(range(numDs)):
            with torch.cuda.stream(streams[i]):
                g_losses[i] = d_branch(i)
        # (join what was forked: the encoder stream only if the DAMSM terms ran on it, here or in the caller)
        for s in streams[:numDs] + ([streams[numDs]] if damsm is not None else []):
            cur.wait_stream(s)
    logs = []
    errG_total = 0
    for i in range(numDs):
        errG_total = errG_total + g_losses[i]
        logs.append(('g_loss%d' % i, g_losses[i].detach()))
    if damsm is not None:
        w_loss, s_loss = damsm
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


def _class_masks(class_ids, batch_size, device):
    """losses.py:24-34 -- mis-match samples of the same class are masked out of the negatives.  ``class_ids`` may already be
    the B x B boolean mask on the device (``class_mask`` below: built once per batch, outside a captured CUDA graph)."""
    if torch.is_tensor(class_ids) and class_ids.dtype == torch.bool and class_ids.dim() == 2:
        return class_ids
    import numpy as np
    ids = np.asarray(class_ids)
    m = (ids[:, None] == ids[None, :])
    m[np.arange(batch_size), np.arange(batch_size)] = False
    return torch.from_numpy(m).to(device)


def class_mask(class_ids, batch_size, device):
    """The B x B mask of ``_class_masks`` as a device tensor (host -> device copy happens here, once per batch)."""
    return _class_masks(class_ids, batch_size, device)


def words_loss(img_features, words_emb, labels, cap_lens, class_ids, batch_size):
    """losses.py:62-132 -- the caption loop + func_attention + cosine + log-sum-exp is ONE fused kernel
    (``mog_damsm_words_fwd/bwd``); only the final B x B cross-entropies are left to torch.  The attention
    maps (third return value, used for visualisation only) are not produced: ``None``."""
    feat = ops.nhwc(img_features)
    B, ih, iw, D = feat.shape
    sims = ops.damsm_similarities(feat.reshape(B, ih * iw, D), words_emb, torch.as_tensor(cap_lens),
                                  cfg.TRAIN.SMOOTH.GAMMA1, cfg.TRAIN.SMOOTH.GAMMA2)
    similarities = sims * cfg.TRAIN.SMOOTH.GAMMA3
    if class_ids is not None:
        similarities = similarities.masked_fill(_class_masks(class_ids, batch_size, similarities.device), -float('inf'))
    if labels is None:
        return None, None, None
    ce = torch.nn.functional.cross_entropy
    return ce(similarities, labels), ce(similarities.transpose(0, 1), labels), None


def sent_loss(cnn_code, rnn_code, labels, class_ids, batch_size, eps=1e-8):
    """losses.py:20-59 -- B x B cosine matrix (GEMM through libmog) * gamma3, two cross-entropies."""
    cnn_code = cnn_code.contiguous()
    rnn_code = rnn_code.detach().contiguous()
    cnn_norm = torch.norm(cnn_code, 2, dim=1, keepdim=True)
    rnn_norm = torch.norm(rnn_code, 2, dim=1, keepdim=True)
    scores0 = ops.linear(cnn_code, rnn_code) / (cnn_norm * rnn_norm.transpose(0, 1)).clamp(min=eps) * cfg.TRAIN.SMOOTH.GAMMA3
    if class_ids is not None:
        scores0 = scores0.masked_fill(_class_masks(class_ids, batch_size, scores0.device), -float('inf'))
    if labels is None:
        return None, None
    ce = torch.nn.functional.cross_entropy
    return ce(scores0, labels), ce(scores0.transpose(0, 1), labels)
