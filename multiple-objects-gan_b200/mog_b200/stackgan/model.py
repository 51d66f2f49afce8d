"""COCO StackGAN-I/II generators and discriminators with the object pathway -- libmog (sm_100a)
edition of the reference's ``code/coco/stackgan/model.py``: same class names, constructor
arguments, ``forward`` signatures, ``cfg`` keys and ``state_dict`` keys/shapes (checked against
``tests/golden/stackgan_keys.json``, dumped from the reference classes).

Same kernel vocabulary as the AttnGAN mirror (``mog_b200/attngan/model.py``): fused
upsample+conv, BN+ReLU/LeakyReLU passes with per-object *segments*, fused STN scatter-sum /
crop(+label concat), NHWC inside, NCHW views outside.  The ``for idx in range(max_objects)`` loops
(model.py:142-147, 212-223, 275-286, 395-405, 414-428, 490-501) are batched object-major; BatchNorm
keeps separate batch statistics and sequential running-stat updates per object, exactly like the
reference's repeated module calls.

Reference quirks kept on purpose: the hard-coded 64 extra fc inputs (model.py:172 => CONDITION_DIM
must be 128 with USE_BBOX_LAYOUT), the hard-coded 128/768 of STAGE2_G (:340, :419 => CONDITION_DIM
128, GF_DIM 192), the frozen STAGE1_G running in train mode inside STAGE2_G (its BatchNorm uses
batch statistics and keeps updating its running stats, :379), ReLU after the residual add (:37-41),
no Sigmoid on the logits (:86-91), 32 -> 31 -> 30 -> 32 in STAGE2_D's object pathway (:466-472, :500).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops
from ..layers import Conv2d, LeakyReLU, ReLU, Tanh, Upsample
from ..ops import ACT_LRELU, ACT_NONE, ACT_RELU, ACT_TANH
from ..stage1_common import D_GET_LOGITS, conv3x3, upBlock  # noqa: F401  (model.py:9-22, 76-104)
from .miscc.config import cfg

N_LABELS = 81


def stn(image, transformation_matrix, size):
    """model.py:107-111 -- single-object form of the fused kernels (kept for API parity)."""
    x = ops.nhwc(image)
    B = x.shape[0]
    y = ops.stn_scatter_sum(x, transformation_matrix.reshape(B, 1, 2, 3), B, 1, (size[2], size[3]), cfg.MOG.ALIGN_CORNERS)
    return ops.to_nchw_view(y)


class ResBlock(nn.Module):
    """model.py:25-41 -- conv, BN, ReLU, conv, BN, += residual, ReLU."""

    def __init__(self, channel_num):
        super().__init__()
        self.block = nn.Sequential(conv3x3(channel_num, channel_num), nn.BatchNorm2d(channel_num), ReLU(True),
                                   conv3x3(channel_num, channel_num), nn.BatchNorm2d(channel_num))
        self.relu = ReLU(inplace=True)

    def forward(self, x):
        b = self.block
        out = ops.bn_act(b[0](x), b[1], ACT_RELU)
        out = ops.bn_act(b[3](out), b[4], ACT_NONE, residual=x)
        return ops.activation(out, ACT_RELU)


class CA_NET(nn.Module):
    """model.py:44-73 -- Linear(+bias) -> ReLU -> (mu | logvar) -> mu + eps * exp(logvar / 2)."""

    def __init__(self):
        super().__init__()
        self.t_dim = cfg.TEXT.DIMENSION
        self.c_dim = cfg.GAN.CONDITION_DIM
        self.fc = nn.Linear(self.t_dim, self.c_dim * 2, bias=True)
        self.relu = ReLU()

    def encode(self, text_embedding):
        x = ops.linear(text_embedding.contiguous(), self.fc.weight, self.fc.bias, act=ACT_RELU)
        return x[:, :self.c_dim], x[:, self.c_dim:]

    def reparametrize(self, mu, logvar, eps=None):
        std = logvar.mul(0.5).exp()
        if eps is None:
            eps = torch.empty_like(std).normal_()   # device generator, like model.py:61
        return eps.mul(std).add(mu)

    def forward(self, text_embedding, eps=None):
        mu, logvar = self.encode(text_embedding)
        return self.reparametrize(mu, logvar, eps), mu, logvar


class BBOX_NET(nn.Module):
    """model.py:114-149 -- label layout (sum of the object labels placed by theta^-1) + three stride-2 convs."""

    def __init__(self):
        super().__init__()
        self.c_dim = cfg.GAN.CONDITION_DIM
        c = self.c_dim
        self.encode = nn.Sequential(
            conv3x3(c, c // 2, stride=2), LeakyReLU(0.2, inplace=True),
            conv3x3(c // 2, c // 4, stride=2), nn.BatchNorm2d(c // 4), LeakyReLU(0.2, inplace=True),
            conv3x3(c // 4, c // 8, stride=2), nn.BatchNorm2d(c // 8), LeakyReLU(0.2, inplace=True))

    def forward_segmajor(self, labels_sb, transf_matr_inv, B, S):
        layout = _label_layout(labels_sb, transf_matr_inv, B, S, self.c_dim)
        e = self.encode
        x = e[0](layout, act=ACT_LRELU)
        x = ops.bn_act(e[2](x), e[3], ACT_LRELU)
        x = ops.bn_act(e[5](x), e[6], ACT_LRELU)
        return x.permute(0, 3, 1, 2).reshape(B, -1)    # the reference flattens NCHW

    def forward(self, labels, transf_matr_inv, max_objects):
        B = labels.shape[0]
        sb = labels[:, :max_objects].transpose(0, 1).reshape(max_objects * B, -1).contiguous()
        return self.forward_segmajor(sb, transf_matr_inv[:, :max_objects].contiguous(), B, max_objects)


def _label_layout(labels_sb, theta_inv, B, S, dim):
    """sum_s stn(label_s replicated over 16x16, theta_inv[:, s]) -> NHWC [B,16,16,dim] (model.py:141-147, 395-405)."""
    planes = labels_sb.reshape(S * B, 1, 1, dim).expand(S * B, 16, 16, dim).contiguous()
    return ops.stn_scatter_sum(planes, theta_inv, B, S, (16, 16), cfg.MOG.ALIGN_CORNERS)


def _object_labels(label_mod, c_code, label_one_hot, S):
    """``self.label(cat(c_code, label_one_hot[:, idx]))`` for idx < S, object-major [S*B, ef]; BatchNorm1d
    statistics per object (model.py:214, 397, 417)."""
    B = c_code.shape[0]
    inp = torch.cat((c_code.unsqueeze(0).expand(S, B, c_code.shape[1]), label_one_hot[:, :S].transpose(0, 1).float()), 2)
    lab = ops.linear(inp.reshape(S * B, -1).contiguous(), label_mod[0].weight)
    return ops.bn_act(lab, label_mod[1], ACT_RELU, segments=S)


# ############# Networks for stageI GAN #############
class STAGE1_G(nn.Module):
    """model.py:152-245"""

    def __init__(self):
        super().__init__()
        self.gf_dim = cfg.GAN.GF_DIM * 8
        self.ef_dim = cfg.GAN.CONDITION_DIM
        self.z_dim = cfg.Z_DIM
        self.define_module()

    def define_module(self):
        ninput = self.z_dim + self.ef_dim
        linput = self.ef_dim + N_LABELS
        ngf = self.gf_dim
        self.ca_net = CA_NET()
        if cfg.USE_BBOX_LAYOUT:
            self.bbox_net = BBOX_NET()
            ninput += 64
        self.fc = nn.Sequential(nn.Linear(ninput, ngf * 4 * 4, bias=False), nn.BatchNorm1d(ngf * 4 * 4), ReLU(True))
        self.label = nn.Sequential(nn.Linear(linput, self.ef_dim, bias=False), nn.BatchNorm1d(self.ef_dim), ReLU(True))
        self.local1 = upBlock(self.ef_dim, ngf // 2)
        self.local2 = upBlock(ngf // 2, ngf // 4)
        self.upsample1 = upBlock(ngf, ngf // 2)
        self.upsample2 = upBlock(ngf // 2, ngf // 4)
        self.upsample3 = upBlock(ngf // 2, ngf // 8)
        self.upsample4 = upBlock(ngf // 8, ngf // 16)
        self.img = nn.Sequential(conv3x3(ngf // 16, 3), Tanh())

    def forward(self, text_embedding, noise, transf_matrices_inv, label_one_hot, max_objects=3, eps=None):
        """``eps``: optional injected CA_NET draw (tests); None = device generator like the reference."""
        S, B = max_objects, noise.shape[0]
        c_code, mu, logvar = self.ca_net(text_embedding, eps)
        tinv = transf_matrices_inv[:, :S].contiguous()
        lab = _object_labels(self.label, c_code, label_one_hot, S)                  # [S*B, ef]
        local_labels = lab.reshape(S, B, self.ef_dim).transpose(0, 1)               # [B, S, ef] (model.py:215)
        h = lab.reshape(S * B, 1, 1, self.ef_dim).expand(S * B, 4, 4, self.ef_dim).contiguous()
        h = self.local1(h, segments=S)
        h = self.local2(h, segments=S)
        h_code_locals = ops.stn_scatter_sum(h, tinv, B, S, (16, 16), cfg.MOG.ALIGN_CORNERS)
        if cfg.USE_BBOX_LAYOUT:
            bbox_code = self.bbox_net.forward_segmajor(lab, tinv, B, S)
            z_c_code = torch.cat((noise, c_code, bbox_code), 1).contiguous()
        else:
            z_c_code = torch.cat((noise, c_code), 1).contiguous()
        h_code = ops.bn_act(ops.linear(z_c_code, self.fc[0].weight), self.fc[1], ACT_RELU)
        h_code = h_code.reshape(B, self.gf_dim, 4, 4).permute(0, 2, 3, 1).contiguous()
        h_code = self.upsample1(h_code)
        h_code = self.upsample2(h_code)
        h_code = torch.cat((h_code, h_code_locals), 3)
        h_code = self.upsample3(h_code)
        h_code = self.upsample4(h_code)
        fake_img = ops.to_nchw_view(self.img[0](h_code, act=ACT_TANH))
        return None, fake_img, mu, logvar, local_labels


class _ObjectPathwayD(nn.Module):
    """Object pathway shared by STAGE1_D / STAGE2_D (model.py:271-286, 487-501): crop the box out of the
    image, concat the one-hot label planes, `local` convs (4x4, stride 1, pad 1: the grid shrinks by one
    per conv), scatter back into the box on an empty canvas."""

    def _locals(self, x, label, transf_matrices, transf_matrices_inv, S, size):
        B = x.shape[0]
        h = ops.stn_crop(x, transf_matrices[:, :S].contiguous(), S, (size, size), extra=label[:, :S].contiguous().float(),
                         align_corners=cfg.MOG.ALIGN_CORNERS)
        loc = self.local
        for i in range(0, len(loc), 3):
            h = ops.bn_act(loc[i](h), loc[i + 1], ACT_LRELU, segments=S)
        return ops.stn_scatter_sum(h, transf_matrices_inv[:, :S].contiguous(), B, S, (size, size), cfg.MOG.ALIGN_CORNERS)


class STAGE1_D(_ObjectPathwayD):
    """model.py:248-309"""

    def __init__(self):
        super().__init__()
        self.df_dim = cfg.GAN.DF_DIM
        self.ef_dim = cfg.GAN.CONDITION_DIM
        self.define_module()

    def define_module(self):
        ndf, nef = self.df_dim, self.ef_dim
        self.local = nn.Sequential(Conv2d(3 + N_LABELS, ndf * 2, 4, 1, 1, bias=False), nn.BatchNorm2d(ndf * 2),
                                   LeakyReLU(0.2, inplace=True))
        self.act = LeakyReLU(0.2, inplace=True)
        self.conv1 = Conv2d(3, ndf, 4, 2, 1, bias=False)
        self.conv2 = Conv2d(ndf, ndf * 2, 4, 2, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(ndf * 2)
        self.conv3 = Conv2d(ndf * 4, ndf * 4, 4, 2, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(ndf * 4)
        self.conv4 = Conv2d(ndf * 4, ndf * 8, 4, 2, 1, bias=False)
        self.bn4 = nn.BatchNorm2d(ndf * 8)
        self.get_cond_logits = D_GET_LOGITS(ndf, nef)
        self.get_uncond_logits = None

    def _encode_img(self, image, label, transf_matrices, transf_matrices_inv, max_objects):
        x = ops.nhwc(image)
        h_code_locals = self._locals(x, label, transf_matrices, transf_matrices_inv, max_objects, 16)
        h = self.conv1(x, act=ACT_LRELU)
        h = ops.bn_act(self.conv2(h), self.bn2, ACT_LRELU)
        h = torch.cat((h, h_code_locals), 3)
        h = ops.bn_act(self.conv3(h), self.bn3, ACT_LRELU)
        h = ops.bn_act(self.conv4(h), self.bn4, ACT_LRELU)
        return ops.to_nchw_view(h)

    def forward(self, image, label, transf_matrices, transf_matrices_inv, max_objects=3):
        return self._encode_img(image, label, transf_matrices, transf_matrices_inv, max_objects)


# ############# Networks for stageII GAN #############
class STAGE2_G(nn.Module):
    """model.py:313-444"""

    def __init__(self, STAGE1_G):
        super().__init__()
        self.gf_dim = cfg.GAN.GF_DIM
        self.ef_dim = cfg.GAN.CONDITION_DIM
        self.z_dim = cfg.Z_DIM
        self.STAGE1_G = STAGE1_G
        for param in self.STAGE1_G.parameters():   # fix parameters of stageI GAN
            param.requires_grad = False
        self.define_module()

    def _make_layer(self, block, channel_num):
        return nn.Sequential(*[block(channel_num) for _ in range(cfg.GAN.R_NUM)])

    def define_module(self):
        ngf = self.gf_dim
        self.ca_net = CA_NET()
        linput = self.ef_dim + N_LABELS
        self.label = nn.Sequential(nn.Linear(linput, self.ef_dim, bias=False), nn.BatchNorm1d(self.ef_dim), ReLU(True))
        self.local1 = upBlock(self.ef_dim + 768, ngf * 2)
        self.local2 = upBlock(ngf * 2, ngf)
        self.encoder = nn.Sequential(
            conv3x3(3, ngf), ReLU(True),
            Conv2d(ngf, ngf * 2, 4, 2, 1, bias=False), nn.BatchNorm2d(ngf * 2), ReLU(True),
            Conv2d(ngf * 2, ngf * 4, 4, 2, 1, bias=False), nn.BatchNorm2d(ngf * 4), ReLU(True))
        joint_in = (self.ef_dim * 2 if cfg.USE_BBOX_LAYOUT else self.ef_dim) + ngf * 4
        self.hr_joint = nn.Sequential(conv3x3(joint_in, ngf * 4), nn.BatchNorm2d(ngf * 4), ReLU(True))
        self.residual = self._make_layer(ResBlock, ngf * 4)
        self.upsample1 = upBlock(ngf * 4, ngf * 2)
        self.upsample2 = upBlock(ngf * 2, ngf)
        self.upsample3 = upBlock(ngf * 2, ngf // 2)
        self.upsample4 = upBlock(ngf // 2, ngf // 4)
        self.img = nn.Sequential(conv3x3(ngf // 4, 3), Tanh())

    def forward(self, text_embedding, noise, transf_matrices_inv, transf_matrices_s2, transf_matrices_inv_s2, label_one_hot,
                max_objects=3, eps=None):
        """``eps``: optional pair (stage-I draw, stage-II draw) of injected CA_NET noise (tests)."""
        S, B = max_objects, noise.shape[0]
        e1, e2 = (None, None) if eps is None else eps
        with torch.no_grad():   # frozen, but in train mode like the reference: batch statistics, running stats updated
            _, stage1_img, _, _, _ = self.STAGE1_G(text_embedding, noise, transf_matrices_inv, label_one_hot, eps=e1)
        stage1_img = stage1_img.detach()
        enc = self.encoder
        x = enc[0](ops.nhwc(stage1_img), act=ACT_RELU)
        x = ops.bn_act(enc[2](x), enc[3], ACT_RELU)
        encoded_img = ops.bn_act(enc[5](x), enc[6], ACT_RELU)                       # [B,16,16,4ngf]

        c_code, mu, logvar = self.ca_net(text_embedding, e2)
        lab = _object_labels(self.label, c_code, label_one_hot, S)                  # [S*B, ef]
        local_labels = lab.reshape(S, B, self.ef_dim).transpose(0, 1)
        c_code_ = c_code.reshape(B, 1, 1, self.ef_dim).expand(B, 16, 16, self.ef_dim)
        if cfg.USE_BBOX_LAYOUT:
            labels_layout = _label_layout(lab, transf_matrices_inv[:, :S].contiguous(), B, S, self.ef_dim)
            i_c_code = torch.cat((encoded_img, c_code_, labels_layout), 3)
        else:
            i_c_code = torch.cat((encoded_img, c_code_), 3)
        h_code = ops.bn_act(self.hr_joint[0](i_c_code), self.hr_joint[1], ACT_RELU)
        for blk in self.residual:
            h_code = blk(h_code)

        # object pathway: crop the box out of the 16x16 feature map, concat the object label, two upBlocks, place at 64x64
        patch = ops.stn_crop(h_code, transf_matrices_s2[:, :S].contiguous(), S, (16, 16), align_corners=cfg.MOG.ALIGN_CORNERS)
        lab_planes = lab.reshape(S * B, 1, 1, self.ef_dim).expand(S * B, 16, 16, self.ef_dim)
        h = torch.cat((patch, lab_planes), 3)
        h = self.local1(h, segments=S)
        h = self.local2(h, segments=S)
        h_code_locals = ops.stn_scatter_sum(h, transf_matrices_inv_s2[:, :S].contiguous(), B, S, (64, 64), cfg.MOG.ALIGN_CORNERS)

        h_code = self.upsample1(h_code)
        h_code = self.upsample2(h_code)
        h_code = torch.cat((h_code, h_code_locals), 3)
        h_code = self.upsample3(h_code)
        h_code = self.upsample4(h_code)
        fake_img = ops.to_nchw_view(self.img[0](h_code, act=ACT_TANH))
        return stage1_img, fake_img, mu, logvar, local_labels


class STAGE2_D(_ObjectPathwayD):
    """model.py:447-537"""

    def __init__(self):
        super().__init__()
        self.df_dim = cfg.GAN.DF_DIM
        self.ef_dim = cfg.GAN.CONDITION_DIM
        self.define_module()

    def define_module(self):
        ndf, nef = self.df_dim, self.ef_dim
        self.local = nn.Sequential(
            Conv2d(3 + N_LABELS, ndf * 2, 4, 1, 1, bias=False), nn.BatchNorm2d(ndf * 2), LeakyReLU(0.2, inplace=True),
            Conv2d(ndf * 2, ndf * 2, 4, 1, 1, bias=False), nn.BatchNorm2d(ndf * 2), LeakyReLU(0.2, inplace=True))
        self.act = LeakyReLU(0.2, inplace=True)
        self.conv1 = Conv2d(3, ndf, 4, 2, 1, bias=False)
        self.conv2 = Conv2d(ndf, ndf * 2, 4, 2, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(ndf * 2)
        self.conv3 = Conv2d(ndf * 2, ndf * 4, 4, 2, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(ndf * 4)
        self.conv4 = Conv2d(ndf * 6, ndf * 8, 4, 2, 1, bias=False)
        self.bn4 = nn.BatchNorm2d(ndf * 8)
        self.conv5 = Conv2d(ndf * 8, ndf * 16, 4, 2, 1, bias=False)
        self.bn5 = nn.BatchNorm2d(ndf * 16)
        self.conv6 = Conv2d(ndf * 16, ndf * 32, 4, 2, 1, bias=False)
        self.bn6 = nn.BatchNorm2d(ndf * 32)
        self.conv7 = conv3x3(ndf * 32, ndf * 16)
        self.bn7 = nn.BatchNorm2d(ndf * 16)
        self.conv8 = conv3x3(ndf * 16, ndf * 8)
        self.bn8 = nn.BatchNorm2d(ndf * 8)
        self.get_cond_logits = D_GET_LOGITS(ndf, nef, bcondition=True)
        self.get_uncond_logits = D_GET_LOGITS(ndf, nef, bcondition=False)

    def _encode_img(self, image, label, transf_matrices, transf_matrices_inv, max_objects):
        x = ops.nhwc(image)
        h_code_locals = self._locals(x, label, transf_matrices, transf_matrices_inv, max_objects, 32)
        h = self.conv1(x, act=ACT_LRELU)
        h = ops.bn_act(self.conv2(h), self.bn2, ACT_LRELU)
        h = ops.bn_act(self.conv3(h), self.bn3, ACT_LRELU)
        h = torch.cat((h, h_code_locals), 3)
        for conv, bn in ((self.conv4, self.bn4), (self.conv5, self.bn5), (self.conv6, self.bn6), (self.conv7, self.bn7),
                         (self.conv8, self.bn8)):
            h = ops.bn_act(conv(h), bn, ACT_LRELU)
        return ops.to_nchw_view(h)

    def forward(self, image, label, transf_matrices, transf_matrices_inv, max_objects=3):
        return self._encode_img(image, label, transf_matrices, transf_matrices_inv, max_objects)
