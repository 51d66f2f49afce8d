"""AttnGAN generator / discriminators with the object pathway -- libmog (sm_100a) edition.

Drop-in for the reference's ``code/coco/attngan/model.py``: same class names, constructor
arguments, ``forward`` signatures, ``cfg`` keys and ``state_dict`` keys/shapes (checked against
``tests/golden/attngan_state_dict_keys.json``, dumped from the reference classes), so
``trainer.py`` / ``train.sh`` and existing checkpoints keep working.  What differs is underneath:
every module runs hand-written CUDA kernels through the C ABI of ``include/mog.h`` on NHWC
activations; the nn.Modules below only own parameters and sequence kernel calls.

Design notes (vs. the reference's per-op torch calls):
* ``nn.Upsample`` + ``conv3x3`` + ``BatchNorm2d`` + ``GLU`` of an upBlock (model.py:48-55) are one
  conv launch with the nearest-neighbour gather folded into the im2col indexing, one statistics
  pass and one fused normalise+GLU pass.
* The ``for idx in range(MAX_OBJECTS)`` loops (model.py:107-112, 393-401, 685-693) are batched:
  the three objects become three *segments* of one 3B-sample launch; BatchNorm keeps separate
  batch statistics (and three sequential running-stat updates) per segment, exactly as three
  separate module calls would.
* ``stn`` (affine_grid + grid_sample) and the canvas ``+=`` / label ``repeat`` / ``cat`` around
  it are single fused scatter-sum / crop+concat kernels.

Public tensors keep the reference's logical NCHW shapes; internally they are contiguous NHWC, so
returned images/features are channels_last views (no copy).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops
from ..ops import ACT_GLU, ACT_LRELU, ACT_NONE, ACT_RELU, ACT_SIGMOID, ACT_TANH
from .GlobalAttention import GlobalAttentionGeneral as ATT_NET
from .miscc.config import cfg

MAX_OBJECTS = 3  # model.py:14


def stn(image, transformation_matrix, size):
    """model.py:17-21 -- ``image`` logical NCHW, ``transformation_matrix`` [B,2,3], ``size`` the
    NCHW output size.  Single-object form of the fused kernels (kept for API parity)."""
    x = ops.nhwc(image)
    B = x.shape[0]
    y = ops.stn_scatter_sum(x, transformation_matrix.reshape(B, 1, 2, 3), B, 1, (size[2], size[3]),
                            cfg.MOG.ALIGN_CORNERS)
    return ops.to_nchw_view(y)


from ..layers import Conv2d, GLU, LeakyReLU, Sigmoid, Tanh, Upsample  # noqa: E402,F401


def conv1x1(in_planes, out_planes, bias=False):
    return Conv2d(in_planes, out_planes, 1, 1, 0, bias=bias)


def conv3x3(in_planes, out_planes, stride=1):
    return Conv2d(in_planes, out_planes, 3, stride, 1, bias=False)


class _UpBlock(nn.Sequential):
    """[Upsample, conv3x3, BatchNorm2d, GLU] -- model.py:48-55."""

    def forward(self, x, segments=1):
        y = self[1](x, up2x=True)
        return ops.bn_act(y, self[2], ACT_GLU, segments=segments)


def upBlock(in_planes, out_planes):
    return _UpBlock(Upsample(scale_factor=2, mode='nearest'), conv3x3(in_planes, out_planes * 2),
                    nn.BatchNorm2d(out_planes * 2), GLU())


class _CBAct(nn.Sequential):
    """[conv, BatchNorm2d, activation marker]"""
    act = ACT_NONE

    def forward(self, x, segments=1):
        return ops.bn_act(self[0](x), self[1], self.act, segments=segments)


class _CBGLU(_CBAct):
    act = ACT_GLU


class _CBLeaky(_CBAct):
    act = ACT_LRELU


def Block3x3_relu(in_planes, out_planes):
    """model.py:59-64 (despite the name: conv, BN, GLU)."""
    return _CBGLU(conv3x3(in_planes, out_planes * 2), nn.BatchNorm2d(out_planes * 2), GLU())


class ResBlock(nn.Module):
    """model.py:67-81 -- conv C->2C, BN, GLU, conv C->C, BN, += residual (fused in the BN pass)."""

    def __init__(self, channel_num):
        super().__init__()
        self.block = nn.Sequential(
            conv3x3(channel_num, channel_num * 2), nn.BatchNorm2d(channel_num * 2), GLU(),
            conv3x3(channel_num, channel_num), nn.BatchNorm2d(channel_num))

    def forward(self, x):
        b = self.block
        out = ops.bn_act(b[0](x), b[1], ACT_GLU)
        return ops.bn_act(b[3](out), b[4], ACT_NONE, residual=x)


class BBOX_NET(nn.Module):
    """model.py:84-116 -- label layout + three stride-2 conv3x3."""

    def __init__(self):
        super().__init__()
        self.c_dim = cfg.GAN.CONDITION_DIM
        c = self.c_dim
        self.encode = nn.Sequential(
            conv3x3(c, c // 2, stride=2), LeakyReLU(0.2, inplace=True),
            conv3x3(c // 2, c // 4, stride=2), nn.BatchNorm2d(c // 4), LeakyReLU(0.2, inplace=True),
            conv3x3(c // 4, c // 8, stride=2), nn.BatchNorm2d(c // 8), LeakyReLU(0.2, inplace=True))

    def forward_segmajor(self, labels_sb, transf_matr_inv, B):
        """labels_sb [S*B, c_dim] (object-major); returns [B, c_dim//8 * 2 * 2]."""
        S = MAX_OBJECTS
        planes = labels_sb.reshape(S * B, 1, 1, self.c_dim).expand(S * B, 16, 16, self.c_dim).contiguous()
        layout = ops.stn_scatter_sum(planes, transf_matr_inv, B, S, (16, 16), cfg.MOG.ALIGN_CORNERS)
        e = self.encode
        x = e[0](layout, act=ACT_LRELU)
        x = ops.bn_act(e[2](x), e[3], ACT_LRELU)
        x = ops.bn_act(e[5](x), e[6], ACT_LRELU)
        # the reference flattens NCHW: (c, h, w) order
        return x.permute(0, 3, 1, 2).reshape(B, -1)

    def forward(self, labels, transf_matr_inv):
        B = labels.shape[0]
        return self.forward_segmajor(labels.transpose(0, 1).reshape(MAX_OBJECTS * B, -1).contiguous(),
                                     transf_matr_inv, B)


# ############## DAMSM image encoder ###################
class RNN_ENCODER(nn.Module):
    """Frozen DAMSM text encoder (model.py:120-204): embedding -> dropout -> bidirectional LSTM / GRU over the packed captions;
    ``words_emb`` [B, nhidden, T] are the per-token outputs, ``sent_emb`` [B, nhidden] the final hidden states of both
    directions.  Same constructor, ``init_hidden`` / ``forward`` signatures and ``state_dict`` keys (``encoder.weight``,
    ``rnn.weight_ih_l0`` ...) as the reference, so ``text_encoder*.pth`` checkpoints load unchanged.  It runs before the
    G/D hot path, once per batch, frozen and in eval mode (trainer.py:78-88): a <= 18-token recurrence with B x 128 states
    is host-side plumbing here (torch's LSTM), not a libmog kernel -- SURVEY.md section 8(f) row f3."""

    def __init__(self, ntoken, ninput=300, drop_prob=0.5, nhidden=128, nlayers=1, bidirectional=True):
        super().__init__()
        self.n_steps = cfg.TEXT.WORDS_NUM
        self.ntoken, self.ninput, self.drop_prob, self.nlayers = ntoken, ninput, drop_prob, nlayers
        self.bidirectional = bidirectional
        self.rnn_type = cfg.RNN_TYPE
        self.num_directions = 2 if bidirectional else 1
        self.nhidden = nhidden // self.num_directions
        self.define_module()
        self.init_weights()

    def define_module(self):
        self.encoder = nn.Embedding(self.ntoken, self.ninput)
        self.drop = nn.Dropout(self.drop_prob)
        if self.rnn_type not in ('LSTM', 'GRU'):
            raise NotImplementedError
        rnn = nn.LSTM if self.rnn_type == 'LSTM' else nn.GRU
        # (dropout between stacked layers only applies for nlayers > 1; torch warns for 1 layer exactly as in the reference)
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            self.rnn = rnn(self.ninput, self.nhidden, self.nlayers, batch_first=True, dropout=self.drop_prob,
                           bidirectional=self.bidirectional)

    def init_weights(self):
        self.encoder.weight.data.uniform_(-0.1, 0.1)

    def init_hidden(self, bsz):
        weight = next(self.parameters())
        z = lambda: weight.new_zeros(self.nlayers * self.num_directions, bsz, self.nhidden)   # noqa: E731
        return (z(), z()) if self.rnn_type == 'LSTM' else z()

    def forward(self, captions, cap_lens, hidden, mask=None):
        from torch.nn.utils.rnn import pack_padded_sequence, pad_packed_sequence
        emb = self.drop(self.encoder(captions))
        lens = cap_lens.detach().cpu().tolist() if torch.is_tensor(cap_lens) else list(cap_lens)
        emb = pack_padded_sequence(emb, lens, batch_first=True)
        output, hidden = self.rnn(emb, hidden)
        output = pad_packed_sequence(output, batch_first=True)[0]
        words_emb = output.transpose(1, 2)
        h = hidden[0] if self.rnn_type == 'LSTM' else hidden
        sent_emb = h.transpose(0, 1).contiguous().view(-1, self.nhidden * self.num_directions)
        return words_emb, sent_emb


class CNN_ENCODER(nn.Module):
    """model.py:207-313 -- frozen Inception-v3 trunk (torchvision layer names, see ``inception.py``) + the two
    trainable-in-DAMSM-pretraining projections ``emb_features`` (1x1 conv 768 -> nef on the 17x17 region map) and
    ``emb_cnn_code`` (Linear 2048 -> nef on the pooled 8x8 map).  The reference downloads the ImageNet weights
    (``model_zoo.load_url``, :215-217); here they are loaded through ``load_state_dict`` (same keys) by the caller --
    there is no network access on the hot path.  ``forward`` -> (region features B x nef x 17 x 17, cnn_code B x nef)."""

    def __init__(self, nef):
        super().__init__()
        self.nef = nef if cfg.TRAIN.FLAG else 256
        from . import inception
        inception.build_trunk(self)
        for param in self.parameters():     # model.py:218-219
            param.requires_grad = False
        self.emb_features = conv1x1(768, self.nef)
        self.emb_cnn_code = nn.Linear(2048, self.nef)
        self.init_trainable_weights()

    def init_trainable_weights(self):
        initrange = 0.1
        self.emb_features.weight.data.uniform_(-initrange, initrange)
        self.emb_cnn_code.weight.data.uniform_(-initrange, initrange)

    def forward(self, x):
        x = ops.resize_bilinear(ops.nhwc(x), (299, 299), align_corners=False)    # :256
        x = self.Conv2d_1a_3x3(x)                 # 149 x 149 x 32
        x = self.Conv2d_2a_3x3(x)                 # 147 x 147 x 32
        x = self.Conv2d_2b_3x3(x)                 # 147 x 147 x 64
        x = ops.max_pool2d(x, 3, 2)               # 73 x 73 x 64
        x = self.Conv2d_3b_1x1(x)                 # 73 x 73 x 80
        x = self.Conv2d_4a_3x3(x)                 # 71 x 71 x 192
        x = ops.max_pool2d(x, 3, 2)               # 35 x 35 x 192
        x = self.Mixed_5d(self.Mixed_5c(self.Mixed_5b(x)))                      # 35 x 35 x 288
        x = self.Mixed_6a(x)                      # 17 x 17 x 768
        x = self.Mixed_6e(self.Mixed_6d(self.Mixed_6c(self.Mixed_6b(x))))
        features = x                              # image region features, 17 x 17 x 768
        x = self.Mixed_7c(self.Mixed_7b(self.Mixed_7a(x)))                      # 8 x 8 x 2048
        x = ops.avg_pool2d(x, 8)                  # 1 x 1 x 2048
        cnn_code = ops.linear(x.reshape(x.shape[0], -1), self.emb_cnn_code.weight, self.emb_cnn_code.bias)
        features = ops.to_nchw_view(self.emb_features(features))
        return features, cnn_code


# ############## G networks ###################
class CA_NET(nn.Module):
    """model.py:317-345"""

    def __init__(self):
        super().__init__()
        self.t_dim = cfg.TEXT.EMBEDDING_DIM
        self.c_dim = cfg.GAN.CONDITION_DIM
        self.fc = nn.Linear(self.t_dim, self.c_dim * 4, bias=True)
        self.relu = GLU()

    def encode(self, text_embedding):
        x = self.relu(ops.linear(text_embedding.contiguous(), self.fc.weight, self.fc.bias))
        return x[:, :self.c_dim], x[:, self.c_dim:]

    def reparametrize(self, mu, logvar, eps=None):
        std = logvar.mul(0.5).exp()
        if eps is None:
            eps = torch.empty_like(std).normal_()  # device generator, like model.py:336
        return eps.mul(std).add(mu)

    def forward(self, text_embedding, eps=None):
        mu, logvar = self.encode(text_embedding)
        return self.reparametrize(mu, logvar, eps), mu, logvar


class INIT_STAGE_G(nn.Module):
    """model.py:348-422"""

    def __init__(self, ngf, ncf):
        super().__init__()
        self.gf_dim = ngf
        self.in_dim = cfg.GAN.Z_DIM + ncf
        self.define_module()

    def define_module(self):
        nz, ngf = self.in_dim, self.gf_dim
        linput = 100 + 81
        self.ef_dim = 100
        self.bbox_net = BBOX_NET()
        nz += 48
        self.fc = nn.Sequential(nn.Linear(nz, ngf * 4 * 4 * 2, bias=False), nn.BatchNorm1d(ngf * 4 * 4 * 2), GLU())
        self.label = nn.Sequential(nn.Linear(linput, self.ef_dim, bias=False), nn.BatchNorm1d(self.ef_dim),
                                   nn.ReLU(True))
        self.local1 = upBlock(self.ef_dim, ngf // 2)
        self.local2 = upBlock(ngf // 2, ngf // 4)
        self.upsample1 = upBlock(ngf, ngf // 2)
        self.upsample2 = upBlock(ngf // 2, ngf // 4)
        self.upsample3 = upBlock(ngf // 2, ngf // 8)
        self.upsample4 = upBlock(ngf // 8, ngf // 16)

    def forward(self, z_code, c_code, transf_matrices_inv, label_one_hot):
        """-> NHWC [B, 64, 64, ngf/16]"""
        B, S = z_code.shape[0], MAX_OBJECTS
        tinv = transf_matrices_inv.contiguous()
        # object pathway, the S objects batched object-major ([S*B, ...]); BN statistics per object
        inp = torch.cat((c_code.unsqueeze(0).expand(S, B, c_code.shape[1]), label_one_hot.transpose(0, 1)), 2)
        lab = ops.linear(inp.reshape(S * B, -1).contiguous(), self.label[0].weight)
        lab = ops.bn_act(lab, self.label[1], ACT_RELU, segments=S)              # [S*B, 100]
        h = lab.reshape(S * B, 1, 1, self.ef_dim).expand(S * B, 4, 4, self.ef_dim).contiguous()
        h = self.local1(h, segments=S)
        h = self.local2(h, segments=S)                                          # [S*B,16,16,ngf/4]
        h_code_locals = ops.stn_scatter_sum(h, tinv, B, S, (16, 16), cfg.MOG.ALIGN_CORNERS)

        bbox_code = self.bbox_net.forward_segmajor(lab, tinv, B)
        c_z_code = torch.cat((c_code, z_code, bbox_code), 1).contiguous()
        out = ops.linear(c_z_code, self.fc[0].weight)
        out = ops.bn_act(out, self.fc[1], ACT_GLU)                               # [B, ngf*16] in (c,h,w) order
        out = out.reshape(B, self.gf_dim, 4, 4).permute(0, 2, 3, 1).contiguous()
        out = self.upsample1(out)
        out = self.upsample2(out)
        out = torch.cat((out, h_code_locals), 3)
        out = self.upsample3(out)
        return self.upsample4(out)


class NEXT_STAGE_G(nn.Module):
    """model.py:425-461"""

    def __init__(self, ngf, nef, ncf):
        super().__init__()
        self.gf_dim, self.ef_dim, self.cf_dim = ngf, nef, ncf
        self.num_residual = cfg.GAN.R_NUM
        self.define_module()

    def _make_layer(self, block, channel_num):
        return nn.Sequential(*[block(channel_num) for _ in range(cfg.GAN.R_NUM)])

    def define_module(self):
        ngf = self.gf_dim
        self.att = ATT_NET(ngf, self.ef_dim)
        self.residual = self._make_layer(ResBlock, ngf * 2)
        self.upsample = upBlock(ngf * 2, ngf)

    def forward(self, h_code, c_code, word_embs, mask):
        """h_code NHWC [B,ih,iw,ngf] -> (NHWC [B,2ih,2iw,ngf], att [B,T,ih,iw])"""
        self.att.applyMask(mask)
        c, att = self.att.forward_nhwc(h_code, word_embs)
        out = torch.cat((h_code, c), 3)
        for blk in self.residual:
            out = blk(out)
        return self.upsample(out), att


class GET_IMAGE_G(nn.Module):
    """model.py:464-475 -- conv3x3 -> 3, tanh (fused epilogue)."""

    def __init__(self, ngf):
        super().__init__()
        self.gf_dim = ngf
        self.img = nn.Sequential(conv3x3(ngf, 3), Tanh())

    def forward(self, h_code):
        return self.img[0](h_code, act=ACT_TANH)


class G_NET(nn.Module):
    """model.py:478-528"""

    def __init__(self):
        super().__init__()
        ngf, nef, ncf = cfg.GAN.GF_DIM, cfg.TEXT.EMBEDDING_DIM, cfg.GAN.CONDITION_DIM
        self.ca_net = CA_NET()
        if cfg.TREE.BRANCH_NUM > 0:
            self.h_net1 = INIT_STAGE_G(ngf * 16, ncf)
            self.img_net1 = GET_IMAGE_G(ngf)
        if cfg.TREE.BRANCH_NUM > 1:
            self.h_net2 = NEXT_STAGE_G(ngf, nef, ncf)
            self.img_net2 = GET_IMAGE_G(ngf)
        if cfg.TREE.BRANCH_NUM > 2:
            self.h_net3 = NEXT_STAGE_G(ngf, nef, ncf)
            self.img_net3 = GET_IMAGE_G(ngf)

    def forward(self, z_code, sent_emb, word_embs, mask, transf_matrices_inv, label_one_hot, eps=None):
        """Same contract as the reference: returns (fake_imgs[NCHW views], att_maps, mu, logvar).
        ``eps`` optionally injects the CA_NET reparametrisation noise (parity tests)."""
        fake_imgs, att_maps = [], []
        c_code, mu, logvar = self.ca_net(sent_emb, eps)
        if cfg.TREE.BRANCH_NUM > 0:
            h_code = self.h_net1(z_code, c_code, transf_matrices_inv, label_one_hot)
            fake_imgs.append(ops.to_nchw_view(self.img_net1(h_code)))
        if cfg.TREE.BRANCH_NUM > 1:
            h_code, att1 = self.h_net2(h_code, c_code, word_embs, mask)
            fake_imgs.append(ops.to_nchw_view(self.img_net2(h_code)))
            if att1 is not None:
                att_maps.append(att1)
        if cfg.TREE.BRANCH_NUM > 2:
            h_code, att2 = self.h_net3(h_code, c_code, word_embs, mask)
            fake_imgs.append(ops.to_nchw_view(self.img_net3(h_code)))
            if att2 is not None:
                att_maps.append(att2)
        return fake_imgs, att_maps, mu, logvar


# ############## D networks ##########################
def Block3x3_leakRelu(in_planes, out_planes):
    """model.py:575-581"""
    return _CBLeaky(conv3x3(in_planes, out_planes), nn.BatchNorm2d(out_planes), LeakyReLU(0.2, inplace=True))


def downBlock(in_planes, out_planes):
    """model.py:585-591"""
    return _CBLeaky(Conv2d(in_planes, out_planes, 4, 2, 1, bias=False), nn.BatchNorm2d(out_planes),
                        LeakyReLU(0.2, inplace=True))


class _Encode16(nn.Sequential):
    """model.py:595-613 (Sequential indices 0..10 as in the reference)."""

    def forward(self, x, segments=1):
        x = self[0](x, act=ACT_LRELU)
        x = ops.bn_act(self[2](x), self[3], ACT_LRELU, segments=segments)
        x = ops.bn_act(self[5](x), self[6], ACT_LRELU, segments=segments)
        return ops.bn_act(self[8](x), self[9], ACT_LRELU, segments=segments)


def encode_image_by_16times(ndf):
    return _Encode16(
        Conv2d(3, ndf, 4, 2, 1, bias=False), LeakyReLU(0.2, inplace=True),
        Conv2d(ndf, ndf * 2, 4, 2, 1, bias=False), nn.BatchNorm2d(ndf * 2), LeakyReLU(0.2, inplace=True),
        Conv2d(ndf * 2, ndf * 4, 4, 2, 1, bias=False), nn.BatchNorm2d(ndf * 4), LeakyReLU(0.2, inplace=True),
        Conv2d(ndf * 4, ndf * 8, 4, 2, 1, bias=False), nn.BatchNorm2d(ndf * 8), LeakyReLU(0.2, inplace=True))


class D_GET_LOGITS(nn.Module):
    """model.py:616-642.  ``forward`` returns sigmoid probabilities like the reference;
    ``logits`` returns the pre-sigmoid values for the fused sigmoid+BCE loss kernel."""

    def __init__(self, ndf, nef, bcondition=False):
        super().__init__()
        self.df_dim, self.ef_dim, self.bcondition = ndf, nef, bcondition
        if self.bcondition:
            self.jointConv = Block3x3_leakRelu(ndf * 8 + nef, ndf * 8)
        self.outlogits = nn.Sequential(Conv2d(ndf * 8, 1, 4, 4, 0, bias=True), Sigmoid())

    def _joint(self, h_code, c_code, segments=1):
        h = ops.nhwc(h_code)
        if self.bcondition and c_code is not None:
            B = h.shape[0]
            c = c_code.reshape(B, 1, 1, self.ef_dim).expand(B, 4, 4, self.ef_dim)
            h = self.jointConv(torch.cat((h, c), 3), segments=segments)
        return h

    def logits(self, h_code, c_code=None, segments=1):
        """``segments`` > 1: ``h_code`` holds that many equally sized batches back to back, each normalised with its
        own BatchNorm statistics -- numerically what consecutive calls on the separate batches do."""
        return self.outlogits[0](self._joint(h_code, c_code, segments)).reshape(-1)

    def forward(self, h_code, c_code=None):
        return self.outlogits[0](self._joint(h_code, c_code), act=ACT_SIGMOID).reshape(-1)


class D_NET64(nn.Module):
    """model.py:646-711 -- with the object pathway."""

    def __init__(self, b_jcu=True):
        super().__init__()
        ndf, nef = cfg.GAN.DF_DIM, cfg.TEXT.EMBEDDING_DIM
        self.UNCOND_DNET = D_GET_LOGITS(ndf, nef, bcondition=False) if b_jcu else None
        self.COND_DNET = D_GET_LOGITS(ndf, nef, bcondition=True)
        self.define_module()

    def define_module(self):
        self.act = LeakyReLU(0.2, inplace=True)
        ndf = cfg.GAN.DF_DIM
        self.conv1 = Conv2d(3, ndf, 4, 2, 1, bias=False)
        self.conv2 = Conv2d(ndf, ndf * 2, 4, 2, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(ndf * 2)
        self.conv3 = Conv2d(ndf * 4, ndf * 4, 4, 2, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(ndf * 4)
        self.conv4 = Conv2d(ndf * 4, ndf * 8, 4, 2, 1, bias=False)
        self.bn4 = nn.BatchNorm2d(ndf * 8)
        self.local = _CBLeaky(Conv2d(3 + 81, ndf * 2, 4, 1, 1, bias=False), nn.BatchNorm2d(ndf * 2),
                                  LeakyReLU(0.2, inplace=True))

    def _locals(self, x, label, transf_matrices, transf_matrices_inv):
        # object pathway: crop each box to 16x16 (+81 label planes), 4x4/s1 conv -> 15x15, BN per
        # object, LeakyReLU, paste back to 16x16 by theta^-1 and sum over objects
        B, S = x.shape[0], MAX_OBJECTS
        h = ops.stn_crop(x, transf_matrices.contiguous(), S, (16, 16), extra=label.contiguous(),
                         align_corners=cfg.MOG.ALIGN_CORNERS)
        h = self.local(h, segments=S)
        return ops.stn_scatter_sum(h, transf_matrices_inv.contiguous(), B, S, (16, 16), cfg.MOG.ALIGN_CORNERS)

    def _trunk(self, x, h_code_locals, segments=1):
        h = self.conv1(x, act=ACT_LRELU)
        h = ops.bn_act(self.conv2(h), self.bn2, ACT_LRELU, segments=segments)
        h = torch.cat((h, h_code_locals), 3)
        h = ops.bn_act(self.conv3(h), self.bn3, ACT_LRELU, segments=segments)
        h = ops.bn_act(self.conv4(h), self.bn4, ACT_LRELU, segments=segments)
        return ops.to_nchw_view(h)

    def forward(self, image, label, transf_matrices, transf_matrices_inv):
        x = ops.nhwc(image)
        return self._trunk(x, self._locals(x, label, transf_matrices, transf_matrices_inv))

    def forward_pair(self, real, fake, label, transf_matrices, transf_matrices_inv):
        """``(self(real, ...), self(fake, ...))`` in one pass: both batches run through every layer back to back as two
        BatchNorm *segments* (own batch statistics, running stats updated real-then-fake like the two calls of
        losses.py:146-152), so each weight is streamed once and its gradient is one reduction over 2B samples.
        The object pathway keeps the reference's order of its six BatchNorm updates (real objects, then fake)."""
        xr, xf = ops.nhwc(real), ops.nhwc(fake)
        loc = torch.cat((self._locals(xr, label, transf_matrices, transf_matrices_inv),
                         self._locals(xf, label, transf_matrices, transf_matrices_inv)), 0)
        return self._trunk(torch.cat((xr, xf), 0), loc, segments=2)


class D_NET128(nn.Module):
    """model.py:715-734"""

    def __init__(self, b_jcu=True):
        super().__init__()
        ndf, nef = cfg.GAN.DF_DIM, cfg.TEXT.EMBEDDING_DIM
        self.img_code_s16 = encode_image_by_16times(ndf)
        self.img_code_s32 = downBlock(ndf * 8, ndf * 16)
        self.img_code_s32_1 = Block3x3_leakRelu(ndf * 16, ndf * 8)
        self.UNCOND_DNET = D_GET_LOGITS(ndf, nef, bcondition=False) if b_jcu else None
        self.COND_DNET = D_GET_LOGITS(ndf, nef, bcondition=True)

    def forward(self, x_var, segments=1):
        x = self.img_code_s16(ops.nhwc(x_var), segments=segments)
        x = self.img_code_s32(x, segments=segments)
        return ops.to_nchw_view(self.img_code_s32_1(x, segments=segments))

    def forward_pair(self, real, fake):
        """Both batches in one pass as two BatchNorm segments (see D_NET64.forward_pair)."""
        return self.forward(torch.cat((ops.nhwc(real), ops.nhwc(fake)), 0).permute(0, 3, 1, 2), segments=2)


class D_NET256(nn.Module):
    """model.py:738-760"""

    def __init__(self, b_jcu=True):
        super().__init__()
        ndf, nef = cfg.GAN.DF_DIM, cfg.TEXT.EMBEDDING_DIM
        self.img_code_s16 = encode_image_by_16times(ndf)
        self.img_code_s32 = downBlock(ndf * 8, ndf * 16)
        self.img_code_s64 = downBlock(ndf * 16, ndf * 32)
        self.img_code_s64_1 = Block3x3_leakRelu(ndf * 32, ndf * 16)
        self.img_code_s64_2 = Block3x3_leakRelu(ndf * 16, ndf * 8)
        self.UNCOND_DNET = D_GET_LOGITS(ndf, nef, bcondition=False) if b_jcu else None
        self.COND_DNET = D_GET_LOGITS(ndf, nef, bcondition=True)

    def forward(self, x_var, segments=1):
        x = self.img_code_s16(ops.nhwc(x_var), segments=segments)
        x = self.img_code_s32(x, segments=segments)
        x = self.img_code_s64(x, segments=segments)
        x = self.img_code_s64_1(x, segments=segments)
        return ops.to_nchw_view(self.img_code_s64_2(x, segments=segments))

    def forward_pair(self, real, fake):
        """Both batches in one pass as two BatchNorm segments (see D_NET64.forward_pair)."""
        return self.forward(torch.cat((ops.nhwc(real), ops.nhwc(fake)), 0).permute(0, 3, 1, 2), segments=2)
