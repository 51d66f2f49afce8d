"""Parameter holders and marker modules shared by the four model programs.

``Conv2d`` keeps nn.Conv2d's ``state_dict`` layout (OIHW weight, optional bias) and runs the libmog
convolution on NHWC activations; the marker classes (``Upsample``, ``LeakyReLU``, ``ReLU``, ``Tanh``,
``Sigmoid``) only occupy the ``nn.Sequential`` slots the reference has, so ``state_dict`` indices
match -- their computation is fused into the neighbouring kernels.  Class names contain 'Conv' /
'BatchNorm' / 'Linear' exactly where the reference's class-name based ``weights_init`` expects them.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops
from .ops import ACT_GLU, ACT_NONE


class Conv2d(nn.Module):
    """Parameter holder with nn.Conv2d's state_dict layout (OIHW weight, optional bias)."""

    def __init__(self, in_planes, out_planes, kernel_size, stride=1, padding=0, bias=False):
        super().__init__()
        self.in_channels, self.out_channels = in_planes, out_planes
        self.kernel_size, self.stride, self.padding = kernel_size, stride, padding
        self.weight = nn.Parameter(torch.empty(out_planes, in_planes, kernel_size, kernel_size))
        self.bias = nn.Parameter(torch.empty(out_planes)) if bias else None
        self.reset_parameters()

    def reset_parameters(self):
        nn.init.kaiming_uniform_(self.weight, a=5 ** 0.5)
        if self.bias is not None:
            fan_in = self.in_channels * self.kernel_size * self.kernel_size
            bound = 1.0 / fan_in ** 0.5
            nn.init.uniform_(self.bias, -bound, bound)

    def forward(self, x, up2x=False, act=ACT_NONE):
        """x NHWC."""
        return ops.conv2d(x, self.weight, self.bias, self.stride, self.padding, up2x, act)

    def extra_repr(self):
        return "%d, %d, kernel_size=%d, stride=%d, padding=%d, bias=%s" % (
            self.in_channels, self.out_channels, self.kernel_size, self.stride, self.padding, self.bias is not None)


class GLU(nn.Module):
    """model.py:24-32 (NHWC: channel halves of the last dim)."""

    def forward(self, x):
        assert x.shape[-1] % 2 == 0, 'channels dont divide 2!'
        return ops.activation(x, ACT_GLU)


class Upsample(nn.Module):
    """Marker for nn.Upsample(scale_factor=2, mode='nearest'): fused into the following conv."""

    def __init__(self, scale_factor=2, mode='nearest'):
        super().__init__()
        self.scale_factor, self.mode = scale_factor, mode


class LeakyReLU(nn.Module):
    """Marker (slope 0.2): fused into the preceding conv epilogue or BN pass."""

    def __init__(self, negative_slope=0.2, inplace=True):
        super().__init__()
        self.negative_slope = negative_slope


class Tanh(nn.Module):
    pass


class Sigmoid(nn.Module):
    pass




class ReLU(nn.Module):
    """Marker: fused into the preceding BN pass."""

    def __init__(self, inplace=True):
        super().__init__()
