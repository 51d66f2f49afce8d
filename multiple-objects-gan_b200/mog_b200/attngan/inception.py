"""Inception-v3 feature trunk of the DAMSM image encoder on libmog kernels (NHWC).

The reference's ``CNN_ENCODER`` (``code/coco/attngan/model.py:207-313``) takes the layers
``Conv2d_1a_3x3 .. Mixed_7c`` of ``torchvision.models.inception_v3()`` (pinned torchvision 0.2.1;
the third-party arithmetic of this path), frozen and in ``eval()`` mode
(``trainer.py:71-77``).  The module tree below reproduces torchvision's names, so ``state_dict``
keys and shapes are identical (``tests/golden/cnn_encoder_keys.json``) and the published
``inception_v3_google`` weights load unchanged.

Every ``BasicConv2d`` (conv without bias -> BatchNorm(eps=1e-3, running statistics) -> ReLU) is ONE
conv launch: the BatchNorm affine is folded into the weights (``w * gamma / sqrt(var + eps)``) and
the bias of the epilogue, ReLU in the epilogue.  Only the data gradient is ever computed (the
encoder is frozen; the gradient w.r.t. the generated image is what ``generator_loss`` needs).
Pooling runs on ``mog_pool2d_*``; branch outputs are concatenated along the NHWC channel axis.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops
from ..ops import ACT_RELU


class _ConvParam(nn.Module):
    """Parameter holder with nn.Conv2d's state_dict layout for a (kh, kw) filter without bias."""

    def __init__(self, cin, cout, kernel_size, stride=1, padding=0):
        super().__init__()
        ks = kernel_size if isinstance(kernel_size, tuple) else (kernel_size, kernel_size)
        pd = padding if isinstance(padding, tuple) else (padding, padding)
        self.in_channels, self.out_channels, self.kernel_size, self.stride, self.padding = cin, cout, ks, stride, pd
        self.weight = nn.Parameter(torch.empty(cout, cin, ks[0], ks[1]))
        nn.init.trunc_normal_(self.weight, std=0.1, a=-0.2, b=0.2)   # torchvision's init (stddev 0.1)


class BasicConv2d(nn.Module):
    """torchvision ``BasicConv2d``: conv(bias=False) + BatchNorm2d(eps=0.001) + ReLU, eval mode only."""

    def __init__(self, cin, cout, **kw):
        super().__init__()
        self.conv = _ConvParam(cin, cout, **kw)
        self.bn = nn.BatchNorm2d(cout, eps=0.001)
        self._folded = None

    def _fold(self):
        srcs = (self.conv.weight, self.bn.weight, self.bn.bias, self.bn.running_mean, self.bn.running_var)
        key = tuple((t.data_ptr(), t._version) for t in srcs)
        if self._folded is None or self._folded[0] != key:
            with torch.no_grad():
                scale = self.bn.weight / torch.sqrt(self.bn.running_var + self.bn.eps)
                w = (self.conv.weight * scale.reshape(-1, 1, 1, 1)).contiguous()
                b = (self.bn.bias - self.bn.running_mean * scale).contiguous()
            self._folded = (key, w, b)
        return self._folded[1], self._folded[2]

    def forward(self, x):
        if self.training:
            raise RuntimeError("the DAMSM image encoder runs frozen in eval() mode (attngan/trainer.py:71-77); "
                               "training-mode BatchNorm of the Inception trunk is not part of the path")
        w, b = self._fold()
        ph, pw = self.conv.padding
        return ops.conv2d(x, w, b, self.conv.stride, ph if ph == pw else (ph, pw), False, ACT_RELU)


def _cat(ts):
    return torch.cat(ts, 3)


class InceptionA(nn.Module):
    def __init__(self, cin, pool_features):
        super().__init__()
        self.branch1x1 = BasicConv2d(cin, 64, kernel_size=1)
        self.branch5x5_1 = BasicConv2d(cin, 48, kernel_size=1)
        self.branch5x5_2 = BasicConv2d(48, 64, kernel_size=5, padding=2)
        self.branch3x3dbl_1 = BasicConv2d(cin, 64, kernel_size=1)
        self.branch3x3dbl_2 = BasicConv2d(64, 96, kernel_size=3, padding=1)
        self.branch3x3dbl_3 = BasicConv2d(96, 96, kernel_size=3, padding=1)
        self.branch_pool = BasicConv2d(cin, pool_features, kernel_size=1)

    def forward(self, x):
        b1 = self.branch1x1(x)
        b5 = self.branch5x5_2(self.branch5x5_1(x))
        b3 = self.branch3x3dbl_3(self.branch3x3dbl_2(self.branch3x3dbl_1(x)))
        bp = self.branch_pool(ops.avg_pool2d(x, 3, 1, 1))
        return _cat((b1, b5, b3, bp))


class InceptionB(nn.Module):
    def __init__(self, cin):
        super().__init__()
        self.branch3x3 = BasicConv2d(cin, 384, kernel_size=3, stride=2)
        self.branch3x3dbl_1 = BasicConv2d(cin, 64, kernel_size=1)
        self.branch3x3dbl_2 = BasicConv2d(64, 96, kernel_size=3, padding=1)
        self.branch3x3dbl_3 = BasicConv2d(96, 96, kernel_size=3, stride=2)

    def forward(self, x):
        b3 = self.branch3x3(x)
        bd = self.branch3x3dbl_3(self.branch3x3dbl_2(self.branch3x3dbl_1(x)))
        return _cat((b3, bd, ops.max_pool2d(x, 3, 2)))


class InceptionC(nn.Module):
    def __init__(self, cin, channels_7x7):
        super().__init__()
        c7 = channels_7x7
        self.branch1x1 = BasicConv2d(cin, 192, kernel_size=1)
        self.branch7x7_1 = BasicConv2d(cin, c7, kernel_size=1)
        self.branch7x7_2 = BasicConv2d(c7, c7, kernel_size=(1, 7), padding=(0, 3))
        self.branch7x7_3 = BasicConv2d(c7, 192, kernel_size=(7, 1), padding=(3, 0))
        self.branch7x7dbl_1 = BasicConv2d(cin, c7, kernel_size=1)
        self.branch7x7dbl_2 = BasicConv2d(c7, c7, kernel_size=(7, 1), padding=(3, 0))
        self.branch7x7dbl_3 = BasicConv2d(c7, c7, kernel_size=(1, 7), padding=(0, 3))
        self.branch7x7dbl_4 = BasicConv2d(c7, c7, kernel_size=(7, 1), padding=(3, 0))
        self.branch7x7dbl_5 = BasicConv2d(c7, 192, kernel_size=(1, 7), padding=(0, 3))
        self.branch_pool = BasicConv2d(cin, 192, kernel_size=1)

    def forward(self, x):
        b1 = self.branch1x1(x)
        b7 = self.branch7x7_3(self.branch7x7_2(self.branch7x7_1(x)))
        bd = self.branch7x7dbl_1(x)
        for m in (self.branch7x7dbl_2, self.branch7x7dbl_3, self.branch7x7dbl_4, self.branch7x7dbl_5):
            bd = m(bd)
        bp = self.branch_pool(ops.avg_pool2d(x, 3, 1, 1))
        return _cat((b1, b7, bd, bp))


class InceptionD(nn.Module):
    def __init__(self, cin):
        super().__init__()
        self.branch3x3_1 = BasicConv2d(cin, 192, kernel_size=1)
        self.branch3x3_2 = BasicConv2d(192, 320, kernel_size=3, stride=2)
        self.branch7x7x3_1 = BasicConv2d(cin, 192, kernel_size=1)
        self.branch7x7x3_2 = BasicConv2d(192, 192, kernel_size=(1, 7), padding=(0, 3))
        self.branch7x7x3_3 = BasicConv2d(192, 192, kernel_size=(7, 1), padding=(3, 0))
        self.branch7x7x3_4 = BasicConv2d(192, 192, kernel_size=3, stride=2)

    def forward(self, x):
        b3 = self.branch3x3_2(self.branch3x3_1(x))
        b7 = self.branch7x7x3_1(x)
        for m in (self.branch7x7x3_2, self.branch7x7x3_3, self.branch7x7x3_4):
            b7 = m(b7)
        return _cat((b3, b7, ops.max_pool2d(x, 3, 2)))


class InceptionE(nn.Module):
    def __init__(self, cin):
        super().__init__()
        self.branch1x1 = BasicConv2d(cin, 320, kernel_size=1)
        self.branch3x3_1 = BasicConv2d(cin, 384, kernel_size=1)
        self.branch3x3_2a = BasicConv2d(384, 384, kernel_size=(1, 3), padding=(0, 1))
        self.branch3x3_2b = BasicConv2d(384, 384, kernel_size=(3, 1), padding=(1, 0))
        self.branch3x3dbl_1 = BasicConv2d(cin, 448, kernel_size=1)
        self.branch3x3dbl_2 = BasicConv2d(448, 384, kernel_size=3, padding=1)
        self.branch3x3dbl_3a = BasicConv2d(384, 384, kernel_size=(1, 3), padding=(0, 1))
        self.branch3x3dbl_3b = BasicConv2d(384, 384, kernel_size=(3, 1), padding=(1, 0))
        self.branch_pool = BasicConv2d(cin, 192, kernel_size=1)

    def forward(self, x):
        b1 = self.branch1x1(x)
        b3 = self.branch3x3_1(x)
        b3 = _cat((self.branch3x3_2a(b3), self.branch3x3_2b(b3)))
        bd = self.branch3x3dbl_2(self.branch3x3dbl_1(x))
        bd = _cat((self.branch3x3dbl_3a(bd), self.branch3x3dbl_3b(bd)))
        bp = self.branch_pool(ops.avg_pool2d(x, 3, 1, 1))
        return _cat((b1, b3, bd, bp))


def build_trunk(owner: nn.Module):
    """Attach the Inception-v3 layers the reference keeps (model.py:227-243) to ``owner`` under torchvision's names."""
    owner.Conv2d_1a_3x3 = BasicConv2d(3, 32, kernel_size=3, stride=2)
    owner.Conv2d_2a_3x3 = BasicConv2d(32, 32, kernel_size=3)
    owner.Conv2d_2b_3x3 = BasicConv2d(32, 64, kernel_size=3, padding=1)
    owner.Conv2d_3b_1x1 = BasicConv2d(64, 80, kernel_size=1)
    owner.Conv2d_4a_3x3 = BasicConv2d(80, 192, kernel_size=3)
    owner.Mixed_5b = InceptionA(192, pool_features=32)
    owner.Mixed_5c = InceptionA(256, pool_features=64)
    owner.Mixed_5d = InceptionA(288, pool_features=64)
    owner.Mixed_6a = InceptionB(288)
    owner.Mixed_6b = InceptionC(768, channels_7x7=128)
    owner.Mixed_6c = InceptionC(768, channels_7x7=160)
    owner.Mixed_6d = InceptionC(768, channels_7x7=160)
    owner.Mixed_6e = InceptionC(768, channels_7x7=192)
    owner.Mixed_7a = InceptionD(768)
    owner.Mixed_7b = InceptionE(1280)
    owner.Mixed_7c = InceptionE(2048)
