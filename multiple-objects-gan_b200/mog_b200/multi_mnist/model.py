"""Multi-MNIST STAGE1_G / STAGE1_D -- libmog edition of ``code/multi-mnist/model.py`` (same class
names, forward signatures, cfg keys and ``state_dict`` layout)."""
from ..stage1_common import (BBOX_NET as _BBOX_NET, D_GET_LOGITS, Flavor, Stage1D, Stage1G, conv3x3,  # noqa: F401
                             upBlock)
from .miscc.config import cfg

FLAVOR = Flavor(n_label=10, img_ch=1, n_objects=3, embed_label=False, bbox_extra=64, bbox_cdim=128, bbox_in=10,
                returns_tuple=True)


class BBOX_NET(_BBOX_NET):
    def __init__(self):
        super().__init__(128, 10)


class STAGE1_G(Stage1G):
    def __init__(self):
        super().__init__(cfg, FLAVOR, ef_dim=10)

    def forward(self, noise, transf_matrices_inv, label_one_hot, num_digits_per_image=3):
        return super().forward(noise, transf_matrices_inv, label_one_hot, num_digits_per_image)


class STAGE1_D(Stage1D):
    def __init__(self):
        super().__init__(cfg, FLAVOR, ef_dim=10)
