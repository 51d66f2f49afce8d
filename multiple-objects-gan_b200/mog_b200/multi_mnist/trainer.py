"""``GANTrainer`` of the Multi-MNIST program -- libmog edition of ``code/multi-mnist/trainer.py`` (training part)."""
from ..attngan.miscc.utils import compute_transformation_matrix, compute_transformation_matrix_inverse
from ..stage1_common import Stage1Trainer
from . import model as _model
from .miscc import utils as _losses
from .miscc.config import cfg as _cfg


class GANTrainer(Stage1Trainer):
    program, cfg, model, losses, n_objects = "mnist", _cfg, _model, _losses, 3

    def unpack_batch(self, data, dev):
        """multi-mnist/trainer.py:114-129 -- (image, bbox [B,3,4], label one-hot [B,3,10]); theta from the boxes."""
        real_img_cpu, bbox, label = data
        real_imgs = real_img_cpu.to(dev, non_blocking=True).float()
        B = real_imgs.shape[0]
        bb = bbox.to(dev).view(-1, 4).float()
        tmi = compute_transformation_matrix_inverse(bb).float().view(B, self.max_objects, 2, 3)
        tm = compute_transformation_matrix(bb).float().view(B, self.max_objects, 2, 3)
        return real_imgs, label.to(dev).float(), tm, tmi
