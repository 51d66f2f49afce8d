"""Losses / init / checkpointing of the COCO StackGAN program -- libmog edition of
``code/coco/stackgan/miscc/utils.py`` (same names and argument order):
``KL_loss`` (:68-71), ``compute_discriminator_loss`` (:74-109, BCEWithLogitsLoss, condition = the
detached mu, "wrong" pairs = shift by one), ``compute_generator_loss`` (:112-125), ``weights_init``
(:129-139, N(0, 0.02) by class name), ``save_model`` (:162-176)."""
from __future__ import annotations

import glob
import os

import torch

from ... import ops
from ...attngan.miscc.utils import (compute_transformation_matrix, compute_transformation_matrix_inverse,  # noqa: F401
                                    mkdir_p)
from ...stage1_common import weights_init  # noqa: F401


def KL_loss(mu, logvar):
    KLD_element = mu.pow(2).add(logvar.exp()).mul(-1).add(1).add(logvar)
    return torch.mean(KLD_element).mul(-0.5)


def _bce(z, t):
    return ops.sigmoid_bce(z, t, with_logits=True)   # nn.BCEWithLogitsLoss


def compute_discriminator_loss(netD, real_imgs, fake_imgs, real_labels, fake_labels, local_label, transf_matrices,
                               transf_matrices_inv, conditions, gpus=None):
    batch_size = real_imgs.size(0)
    cond = conditions.detach()
    fake = fake_imgs.detach()
    local_label = local_label.detach()
    real_features = netD(real_imgs, local_label, transf_matrices, transf_matrices_inv)
    fake_features = netD(fake, local_label, transf_matrices, transf_matrices_inv)
    errD_real = _bce(netD.get_cond_logits(real_features, cond), real_labels)
    errD_wrong = _bce(netD.get_cond_logits(real_features[:(batch_size - 1)], cond[1:]), fake_labels[1:])
    errD_fake = _bce(netD.get_cond_logits(fake_features, cond), fake_labels)
    if netD.get_uncond_logits is not None:
        uncond_errD_real = _bce(netD.get_uncond_logits(real_features), real_labels)
        uncond_errD_fake = _bce(netD.get_uncond_logits(fake_features), fake_labels)
        errD = ((errD_real + uncond_errD_real) / 2. + (errD_fake + errD_wrong + uncond_errD_fake) / 3.)
        errD_real = (errD_real + uncond_errD_real) / 2.
        errD_fake = (errD_fake + uncond_errD_fake) / 2.
    else:
        errD = errD_real + (errD_fake + errD_wrong) * 0.5
    # the reference returns .item() floats for the three parts (3 host syncs per step); device scalars here
    return errD, errD_real.detach(), errD_wrong.detach(), errD_fake.detach()


def compute_generator_loss(netD, fake_imgs, real_labels, local_label, transf_matrices, transf_matrices_inv, conditions,
                           gpus=None):
    cond = conditions.detach()
    fake_features = netD(fake_imgs, local_label, transf_matrices, transf_matrices_inv)
    errD_fake = _bce(netD.get_cond_logits(fake_features, cond), real_labels)
    if netD.get_uncond_logits is not None:
        errD_fake = errD_fake + _bce(netD.get_uncond_logits(fake_features), real_labels)
    return errD_fake


def save_model(netG, netD, optimG, optimD, epoch, model_dir, saveD=False, saveOptim=False, max_to_keep=5):
    """Checkpoint wire format of the reference (:162-176): ``{epoch, netG, optimG, netD, optimD}``."""
    checkpoint = {'epoch': epoch, 'netG': netG.state_dict(), 'optimG': optimG.state_dict() if saveOptim else {},
                  'netD': netD.state_dict() if saveD else {}, 'optimD': optimD.state_dict() if saveOptim else {}}
    torch.save(checkpoint, "{}/checkpoint_{:04}.pth".format(model_dir, epoch))
    if max_to_keep is not None and max_to_keep > 0:
        checkpoint_list = sorted(glob.glob(model_dir + "/" + '*.pth'))
        while len(checkpoint_list) > max_to_keep:
            os.remove(checkpoint_list[0])
            checkpoint_list = checkpoint_list[1:]
