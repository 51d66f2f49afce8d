"""CLEVR STAGE1_G / STAGE1_D -- libmog edition of ``code/clevr/model.py`` (64x64, 4 object slots,
13-dim labels = one-hot shape(4) + colour(9); empty slots are all -1 rows)."""
from ..stage1_common import (BBOX_NET as _BBOX_NET, D_GET_LOGITS, Flavor, Stage1D, Stage1G, conv3x3,  # noqa: F401
                             upBlock)
from .miscc.config import cfg


def _flavor():
    c = cfg.GAN.CONDITION_DIM
    return Flavor(n_label=13, img_ch=3, n_objects=4, embed_label=True, bbox_extra=8, bbox_cdim=c, bbox_in=c,
                  returns_tuple=False)


class BBOX_NET(_BBOX_NET):
    def __init__(self):
        super().__init__(cfg.GAN.CONDITION_DIM, cfg.GAN.CONDITION_DIM)


class STAGE1_G(Stage1G):
    def __init__(self):
        super().__init__(cfg, _flavor(), ef_dim=cfg.GAN.CONDITION_DIM)

    def forward(self, noise, transf_matrices_inv, label_one_hot, num_objects=4):
        return super().forward(noise, transf_matrices_inv, label_one_hot, num_objects)


class STAGE1_D(Stage1D):
    def __init__(self):
        super().__init__(cfg, _flavor(), ef_dim=cfg.GAN.CONDITION_DIM)
