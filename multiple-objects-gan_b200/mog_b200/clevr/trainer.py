"""``GANTrainer`` of the CLEVR program -- libmog edition of ``code/clevr/trainer.py`` (training part)."""
from ..stage1_common import Stage1Trainer
from . import model as _model
from .miscc import utils as _losses
from .miscc.config import cfg as _cfg


class GANTrainer(Stage1Trainer):
    program, cfg, model, losses, n_objects = "clevr", _cfg, _model, _losses, 4

    def unpack_batch(self, data, dev):
        """clevr/trainer.py:114-125 -- (image, [theta, theta^-1], label one-hot [B,4,13], _)."""
        real_img_cpu, transformation_matrices, label_one_hot = data[0], data[1], data[2]
        tm, tmi = tuple(transformation_matrices)
        return (real_img_cpu.to(dev, non_blocking=True).float(), label_one_hot.to(dev).float(), tm.detach().to(dev).float(),
                tmi.detach().to(dev).float())
