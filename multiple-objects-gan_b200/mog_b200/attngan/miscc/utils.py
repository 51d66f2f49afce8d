"""Host-side helpers of the AttnGAN program with the reference's names and semantics
(``code/coco/attngan/miscc/utils.py``): bbox -> affine matrices (lines 16-49), the class-name
based ``weights_init`` (321-331) and the EMA parameter copies (334-341).  Visualisation helpers
of the reference (51-317) are host-side plotting and out of scope."""
from __future__ import annotations

import errno
import os
from copy import deepcopy

import torch
import torch.nn as nn


def compute_transformation_matrix_inverse(bbox):
    """(x, y, w, h) fractions -> theta^-1 [N,2,3] placing a full canvas into the box."""
    x, y, w, h = bbox[:, 0], bbox[:, 1], bbox[:, 2], bbox[:, 3]
    scale_x, scale_y = 1.0 / w, 1.0 / h
    t_x = 2 * scale_x * (0.5 - (x + 0.5 * w))
    t_y = 2 * scale_y * (0.5 - (y + 0.5 * h))
    zeros = torch.zeros_like(x)
    return torch.stack([scale_x, zeros, t_x, zeros, scale_y, t_y], 1).view(-1, 2, 3)


def compute_transformation_matrix(bbox):
    """(x, y, w, h) fractions -> theta [N,2,3] cropping the box out of the image."""
    x, y, w, h = bbox[:, 0], bbox[:, 1], bbox[:, 2], bbox[:, 3]
    t_x = 2 * ((x + 0.5 * w) - 0.5)
    t_y = 2 * ((y + 0.5 * h) - 0.5)
    zeros = torch.zeros_like(x)
    return torch.stack([w, zeros, t_x, zeros, h, t_y], 1).view(-1, 2, 3)


def weights_init(m):
    from ... import ops
    ops.invalidate_packed(m.parameters(recurse=False))     # .data writes below are invisible to torch's version counter
    classname = m.__class__.__name__
    if classname.find('Conv') != -1:
        nn.init.orthogonal_(m.weight.data, 1.0)
    elif classname.find('BatchNorm') != -1:
        m.weight.data.normal_(1.0, 0.02)
        m.bias.data.fill_(0)
    elif classname.find('Linear') != -1:
        nn.init.orthogonal_(m.weight.data, 1.0)
        if m.bias is not None:
            m.bias.data.fill_(0.0)


def load_params(model, new_param):
    from ... import ops
    for p, new_p in zip(model.parameters(), new_param):
        p.data.copy_(new_p)
    ops.invalidate_packed(model)      # the packed conv operands of every weight are stale now


def copy_G_params(model):
    return deepcopy(list(p.data for p in model.parameters()))


def mkdir_p(path):
    try:
        os.makedirs(path)
    except OSError as exc:
        if exc.errno == errno.EEXIST and os.path.isdir(path):
            pass
        else:
            raise
