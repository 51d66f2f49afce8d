"""Losses / init of the CLEVR program (``code/clevr/miscc/utils.py:93-142``)."""
from ...attngan.miscc.utils import compute_transformation_matrix, compute_transformation_matrix_inverse  # noqa: F401
from ...stage1_common import weights_init  # noqa: F401
from ... import stage1_common as _c


def compute_discriminator_loss(netD, real_imgs, fake_imgs, real_labels, fake_labels, local_label, transf_matrices,
                               transf_matrices_inv, gpus=None):
    return _c.compute_discriminator_loss(netD, real_imgs, fake_imgs, real_labels, fake_labels, local_label,
                                         transf_matrices, transf_matrices_inv, gpus, n_objects=4, clamp_negative=True)


def compute_generator_loss(netD, fake_imgs, real_labels, local_label, transf_matrices, transf_matrices_inv, gpus=None):
    return _c.compute_generator_loss(netD, fake_imgs, real_labels, local_label, transf_matrices, transf_matrices_inv,
                                     gpus, n_objects=4, clamp_negative=True)
