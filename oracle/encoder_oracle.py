"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product path.

CPU fp32 restatement of the reference's DAMSM image encoder ``CNN_ENCODER``
(``code/coco/attngan/model.py:207-313``) over a ``state_dict``.  The arithmetic of the trunk lives in a
third-party dependency that is NOT vendored in /root/reference: torchvision's ``inception_v3``
(pinned ``torchvision==0.2.1``, requirements.txt:31; the container has 0.26).  Its published
algorithm (Szegedy et al., "Rethinking the Inception Architecture", blocks A-E as implemented in
``torchvision/models/inception.py``) is restated here: BasicConv2d = conv(bias=False) +
BatchNorm(eps=1e-3) + ReLU; max pools 3x3/2; branch average pools 3x3/1 pad 1 (count_include_pad).
Pinned against vectors produced by executing the unmodified reference class on top of the
container's torchvision (``tests/golden/make_golden_encoder.py`` -> ``tests/golden/cnn_encoder.npz``)
by ``tests/test_encoder.py``.  Eval mode only (the reference freezes the encoder, trainer.py:71-77).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def bconv(x, P, pre, stride=1, padding=0):
    """torchvision BasicConv2d in eval mode."""
    x = F.conv2d(x, P[pre + ".conv.weight"], None, stride, padding)
    x = F.batch_norm(x, P[pre + ".bn.running_mean"], P[pre + ".bn.running_var"], P[pre + ".bn.weight"], P[pre + ".bn.bias"],
                     training=False, eps=0.001)
    return F.relu(x)


def inception_a(x, P, p):
    b1 = bconv(x, P, p + ".branch1x1")
    b5 = bconv(bconv(x, P, p + ".branch5x5_1"), P, p + ".branch5x5_2", padding=2)
    b3 = bconv(x, P, p + ".branch3x3dbl_1")
    b3 = bconv(b3, P, p + ".branch3x3dbl_2", padding=1)
    b3 = bconv(b3, P, p + ".branch3x3dbl_3", padding=1)
    bp = bconv(F.avg_pool2d(x, 3, 1, 1), P, p + ".branch_pool")
    return torch.cat((b1, b5, b3, bp), 1)


def inception_b(x, P, p):
    b3 = bconv(x, P, p + ".branch3x3", stride=2)
    bd = bconv(x, P, p + ".branch3x3dbl_1")
    bd = bconv(bd, P, p + ".branch3x3dbl_2", padding=1)
    bd = bconv(bd, P, p + ".branch3x3dbl_3", stride=2)
    return torch.cat((b3, bd, F.max_pool2d(x, 3, 2)), 1)


def inception_c(x, P, p):
    b1 = bconv(x, P, p + ".branch1x1")
    b7 = bconv(x, P, p + ".branch7x7_1")
    b7 = bconv(b7, P, p + ".branch7x7_2", padding=(0, 3))
    b7 = bconv(b7, P, p + ".branch7x7_3", padding=(3, 0))
    bd = bconv(x, P, p + ".branch7x7dbl_1")
    bd = bconv(bd, P, p + ".branch7x7dbl_2", padding=(3, 0))
    bd = bconv(bd, P, p + ".branch7x7dbl_3", padding=(0, 3))
    bd = bconv(bd, P, p + ".branch7x7dbl_4", padding=(3, 0))
    bd = bconv(bd, P, p + ".branch7x7dbl_5", padding=(0, 3))
    bp = bconv(F.avg_pool2d(x, 3, 1, 1), P, p + ".branch_pool")
    return torch.cat((b1, b7, bd, bp), 1)


def inception_d(x, P, p):
    b3 = bconv(bconv(x, P, p + ".branch3x3_1"), P, p + ".branch3x3_2", stride=2)
    b7 = bconv(x, P, p + ".branch7x7x3_1")
    b7 = bconv(b7, P, p + ".branch7x7x3_2", padding=(0, 3))
    b7 = bconv(b7, P, p + ".branch7x7x3_3", padding=(3, 0))
    b7 = bconv(b7, P, p + ".branch7x7x3_4", stride=2)
    return torch.cat((b3, b7, F.max_pool2d(x, 3, 2)), 1)


def inception_e(x, P, p):
    b1 = bconv(x, P, p + ".branch1x1")
    b3 = bconv(x, P, p + ".branch3x3_1")
    b3 = torch.cat((bconv(b3, P, p + ".branch3x3_2a", padding=(0, 1)), bconv(b3, P, p + ".branch3x3_2b", padding=(1, 0))), 1)
    bd = bconv(bconv(x, P, p + ".branch3x3dbl_1"), P, p + ".branch3x3dbl_2", padding=1)
    bd = torch.cat((bconv(bd, P, p + ".branch3x3dbl_3a", padding=(0, 1)), bconv(bd, P, p + ".branch3x3dbl_3b", padding=(1, 0))), 1)
    bp = bconv(F.avg_pool2d(x, 3, 1, 1), P, p + ".branch_pool")
    return torch.cat((b1, b3, bd, bp), 1)


def cnn_encoder(P, x):
    """model.py:252-313 -> (region features B x nef x 17 x 17, cnn_code B x nef)."""
    x = F.interpolate(x, size=(299, 299), mode="bilinear", align_corners=False)   # nn.Upsample(size, 'bilinear'), :256
    x = bconv(x, P, "Conv2d_1a_3x3", stride=2)
    x = bconv(x, P, "Conv2d_2a_3x3")
    x = bconv(x, P, "Conv2d_2b_3x3", padding=1)
    x = F.max_pool2d(x, 3, 2)
    x = bconv(x, P, "Conv2d_3b_1x1")
    x = bconv(x, P, "Conv2d_4a_3x3")
    x = F.max_pool2d(x, 3, 2)
    for n in ("Mixed_5b", "Mixed_5c", "Mixed_5d"):
        x = inception_a(x, P, n)
    x = inception_b(x, P, "Mixed_6a")
    for n in ("Mixed_6b", "Mixed_6c", "Mixed_6d", "Mixed_6e"):
        x = inception_c(x, P, n)
    features = x
    x = inception_d(x, P, "Mixed_7a")
    x = inception_e(x, P, "Mixed_7b")
    x = inception_e(x, P, "Mixed_7c")
    x = F.avg_pool2d(x, 8).reshape(x.shape[0], -1)
    cnn_code = F.linear(x, P["emb_cnn_code.weight"], P["emb_cnn_code.bias"])
    return F.conv2d(features, P["emb_features.weight"]), cnn_code
