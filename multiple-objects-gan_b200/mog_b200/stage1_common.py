"""Shared implementation of the single-stage (64x64) programs of the reference: Multi-MNIST
(``code/multi-mnist/model.py``) and CLEVR (``code/clevr/model.py``).  The two files differ only in
label width (10 / 13), image channels (1 / 3), number of objects (3 / 4) and in whether the object
label goes through a Linear+BN1d+ReLU embedding (CLEVR, ``clevr/model.py:137-140,164``) -- see the
diff in SURVEY.md section 8 a22.  The concrete ``STAGE1_G`` / ``STAGE1_D`` classes live in
``mog_b200/multi_mnist/model.py`` and ``mog_b200/clevr/model.py`` and bind a ``Flavor``.

Same kernel vocabulary as the AttnGAN mirror: fused upsample+conv, BN+ReLU/LeakyReLU passes with
per-object segments, fused STN scatter-sum / crop+label-concat, NHWC inside, NCHW views outside.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.nn as nn

from . import ops
from .layers import Conv2d, LeakyReLU, ReLU, Tanh, Upsample
from .ops import ACT_LRELU, ACT_NONE, ACT_RELU, ACT_TANH


@dataclass
class Flavor:
    n_label: int          # width of the one-hot label (10 MNIST, 13 CLEVR)
    img_ch: int           # image channels
    n_objects: int        # objects per image
    embed_label: bool     # CLEVR: label -> Linear+BN1d+ReLU -> ef_dim
    bbox_extra: int       # features the bbox encoder adds to the fc input (64 MNIST, 8 CLEVR)
    bbox_cdim: int        # BBOX_NET.c_dim (128 MNIST, cfg.GAN.CONDITION_DIM CLEVR)
    bbox_in: int          # channels of the label layout (10 MNIST, c_dim CLEVR)
    returns_tuple: bool   # MNIST returns (None, img)


def conv3x3(in_planes, out_planes, stride=1):
    return Conv2d(in_planes, out_planes, 3, stride, 1, bias=False)


class _UpBlockReLU(nn.Sequential):
    """[Upsample, conv3x3, BatchNorm2d, ReLU] -- multi-mnist/model.py:16-22"""

    def forward(self, x, segments=1):
        return ops.bn_act(self[1](x, up2x=True), self[2], ACT_RELU, segments=segments)


def upBlock(in_planes, out_planes):
    return _UpBlockReLU(Upsample(scale_factor=2, mode='nearest'), conv3x3(in_planes, out_planes),
                        nn.BatchNorm2d(out_planes), ReLU(True))


class D_GET_LOGITS(nn.Module):
    """multi-mnist/model.py:44-71, clevr/model.py:44-71 -- no Sigmoid (BCEWithLogitsLoss)."""

    def __init__(self, ndf, nef, bcondition=True):
        super().__init__()
        self.df_dim, self.ef_dim, self.bcondition = ndf, nef, bcondition
        if bcondition:
            self.outlogits = nn.Sequential(conv3x3(ndf * 8 + nef, ndf * 8), nn.BatchNorm2d(ndf * 8),
                                           LeakyReLU(0.2, inplace=True), Conv2d(ndf * 8, 1, 4, 4, 0, bias=True))
        else:
            self.outlogits = nn.Sequential(Conv2d(ndf * 8, 1, 4, 4, 0, bias=True))

    def forward(self, h_code, c_code=None):
        h = ops.nhwc(h_code)
        o = self.outlogits
        if self.bcondition and c_code is not None:
            B = h.shape[0]
            c = c_code.reshape(B, 1, 1, self.ef_dim).expand(B, 4, 4, self.ef_dim)
            h = ops.bn_act(o[0](torch.cat((h, c), 3)), o[1], ACT_LRELU)
            return o[3](h).reshape(-1)
        return o[0](h).reshape(-1)


class BBOX_NET(nn.Module):
    """multi-mnist/model.py:81-111 / clevr/model.py:81-111"""

    def __init__(self, c_dim, in_ch):
        super().__init__()
        self.c_dim, self.in_ch = c_dim, in_ch
        c = c_dim
        self.encode = nn.Sequential(
            conv3x3(in_ch, c // 2, stride=2), LeakyReLU(0.2, inplace=True),
            conv3x3(c // 2, c // 4, stride=2), nn.BatchNorm2d(c // 4), LeakyReLU(0.2, inplace=True),
            conv3x3(c // 4, c // 8, stride=2), nn.BatchNorm2d(c // 8), LeakyReLU(0.2, inplace=True))

    def forward_segmajor(self, labels_sb, transf_matr_inv, B, S):
        planes = labels_sb.reshape(S * B, 1, 1, self.in_ch).expand(S * B, 16, 16, self.in_ch).contiguous()
        layout = ops.stn_scatter_sum(planes, transf_matr_inv, B, S, (16, 16))
        e = self.encode
        x = e[0](layout, act=ACT_LRELU)
        x = ops.bn_act(e[2](x), e[3], ACT_LRELU)
        x = ops.bn_act(e[5](x), e[6], ACT_LRELU)
        return x.permute(0, 3, 1, 2).reshape(B, -1)

    def forward(self, labels, transf_matr_inv, num_objects):
        B = labels.shape[0]
        sb = labels[:, :num_objects].transpose(0, 1).reshape(num_objects * B, -1).contiguous()
        return self.forward_segmajor(sb, transf_matr_inv[:, :num_objects].contiguous(), B, num_objects)


class Stage1G(nn.Module):
    """multi-mnist/model.py:113-192, clevr/model.py:113-194"""

    def __init__(self, cfg, fl: Flavor, ef_dim):
        super().__init__()
        self.fl = fl
        self.gf_dim = cfg.GAN.GF_DIM * 8
        self.ef_dim = ef_dim
        self.z_dim = cfg.Z_DIM
        self.use_bbox = bool(cfg.USE_BBOX_LAYOUT)
        ninput, ngf = self.z_dim, self.gf_dim
        if self.use_bbox:
            self.bbox_net = BBOX_NET(fl.bbox_cdim, fl.bbox_in)
            ninput += fl.bbox_extra
        self.fc = nn.Sequential(nn.Linear(ninput, ngf * 4 * 4, bias=False), nn.BatchNorm1d(ngf * 4 * 4), ReLU(True))
        self.label = nn.Sequential(nn.Linear(fl.n_label, self.ef_dim, bias=False), nn.BatchNorm1d(self.ef_dim), ReLU(True))
        self.local1 = upBlock(self.ef_dim, ngf // 2)
        self.local2 = upBlock(ngf // 2, ngf // 4)
        self.upsample1 = upBlock(ngf, ngf // 2)
        self.upsample2 = upBlock(ngf // 2, ngf // 4)
        self.upsample3 = upBlock(ngf // 2, ngf // 8)
        self.upsample4 = upBlock(ngf // 8, ngf // 16)
        self.img = nn.Sequential(conv3x3(ngf // 16, fl.img_ch), Tanh())

    def forward(self, noise, transf_matrices_inv, label_one_hot, num_objects=None):
        S = self.fl.n_objects if num_objects is None else num_objects
        B = noise.shape[0]
        tinv = transf_matrices_inv[:, :S].contiguous()
        lab = label_one_hot[:, :S].transpose(0, 1).reshape(S * B, -1).contiguous().float()   # object-major
        if self.fl.embed_label:
            lab = ops.bn_act(ops.linear(lab, self.label[0].weight), self.label[1], ACT_RELU, segments=S)
        h = lab.reshape(S * B, 1, 1, self.ef_dim).expand(S * B, 4, 4, self.ef_dim).contiguous()
        h = self.local1(h, segments=S)
        h = self.local2(h, segments=S)
        h_code_locals = ops.stn_scatter_sum(h, tinv, B, S, (16, 16))
        if self.use_bbox:
            bbox_code = self.bbox_net.forward_segmajor(lab, tinv, B, S)
            z_c_code = torch.cat((noise, bbox_code), 1).contiguous()
        else:
            z_c_code = noise.contiguous()
        h_code = ops.bn_act(ops.linear(z_c_code, self.fc[0].weight), self.fc[1], ACT_RELU)
        h_code = h_code.reshape(B, self.gf_dim, 4, 4).permute(0, 2, 3, 1).contiguous()
        h_code = self.upsample1(h_code)
        h_code = self.upsample2(h_code)
        h_code = torch.cat((h_code, h_code_locals), 3)
        h_code = self.upsample3(h_code)
        h_code = self.upsample4(h_code)
        fake_img = ops.to_nchw_view(self.img[0](h_code, act=ACT_TANH))
        return (None, fake_img) if self.fl.returns_tuple else fake_img


class Stage1D(nn.Module):
    """multi-mnist/model.py:195-257, clevr/model.py:197-260"""

    def __init__(self, cfg, fl: Flavor, ef_dim):
        super().__init__()
        self.fl = fl
        self.df_dim = cfg.GAN.DF_DIM
        self.ef_dim = ef_dim
        ndf = self.df_dim
        self.local = nn.Sequential(Conv2d(fl.img_ch + fl.n_label, ndf * 2, 4, 1, 1, bias=False), nn.BatchNorm2d(ndf * 2),
                                   LeakyReLU(0.2, inplace=True))
        self.act = LeakyReLU(0.2, inplace=True)
        self.conv1 = Conv2d(fl.img_ch, ndf, 4, 2, 1, bias=False)
        self.conv2 = Conv2d(ndf, ndf * 2, 4, 2, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(ndf * 2)
        self.conv3 = Conv2d(ndf * 4, ndf * 4, 4, 2, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(ndf * 4)
        self.conv4 = Conv2d(ndf * 4, ndf * 8, 4, 2, 1, bias=False)
        self.bn4 = nn.BatchNorm2d(ndf * 8)
        self.get_cond_logits = D_GET_LOGITS(ndf, fl.n_label)
        self.get_uncond_logits = None

    def _encode_img(self, image, label, transf_matrices, transf_matrices_inv, num_objects=None):
        S = self.fl.n_objects if num_objects is None else num_objects
        x = ops.nhwc(image)
        B = x.shape[0]
        h = ops.stn_crop(x, transf_matrices[:, :S].contiguous(), S, (16, 16), extra=label[:, :S].contiguous().float())
        h = ops.bn_act(self.local[0](h), self.local[1], ACT_LRELU, segments=S)
        h_code_locals = ops.stn_scatter_sum(h, transf_matrices_inv[:, :S].contiguous(), B, S, (16, 16))
        h = self.conv1(x, act=ACT_LRELU)
        h = ops.bn_act(self.conv2(h), self.bn2, ACT_LRELU)
        h = torch.cat((h, h_code_locals), 3)
        h = ops.bn_act(self.conv3(h), self.bn3, ACT_LRELU)
        h = ops.bn_act(self.conv4(h), self.bn4, ACT_LRELU)
        return ops.to_nchw_view(h)

    def forward(self, image, label, transf_matrices, transf_matrices_inv):
        return self._encode_img(image, label, transf_matrices, transf_matrices_inv)


# ---- losses (multi-mnist/miscc/utils.py:71-123, clevr/miscc/utils.py:93-142) ---------------------
def _label_cond(local_label, n_objects, clamp_negative):
    cond = local_label[:, 0, :].clone()
    for i in range(1, n_objects):
        cond = cond + local_label[:, i, :]
    if clamp_negative:   # clevr/miscc/utils.py:99: empty slots are encoded as -1 rows
        cond = cond.clamp_min(0)
    return cond.float()


def compute_discriminator_loss(netD, real_imgs, fake_imgs, real_labels, fake_labels, local_label, transf_matrices,
                               transf_matrices_inv, gpus=None, n_objects=3, clamp_negative=False):
    bce = lambda z, t: ops.sigmoid_bce(z, t, with_logits=True)   # nn.BCEWithLogitsLoss
    batch_size = real_imgs.size(0)
    fake = fake_imgs.detach()
    local_label = local_label.detach()
    cond = _label_cond(local_label, n_objects, clamp_negative)
    real_features = netD(real_imgs, local_label, transf_matrices, transf_matrices_inv)
    fake_features = netD(fake, local_label, transf_matrices, transf_matrices_inv)
    errD_real = bce(netD.get_cond_logits(real_features, cond), real_labels)
    errD_wrong = bce(netD.get_cond_logits(real_features[:(batch_size - 1)], cond[1:]), fake_labels[1:])
    errD_fake = bce(netD.get_cond_logits(fake_features, cond), fake_labels)
    if netD.get_uncond_logits is not None:
        uncond_real = bce(netD.get_uncond_logits(real_features), real_labels)
        uncond_fake = bce(netD.get_uncond_logits(fake_features), fake_labels)
        errD = ((errD_real + uncond_real) / 2. + (errD_fake + errD_wrong + uncond_fake) / 3.)
        errD_real = (errD_real + uncond_real) / 2.
        errD_fake = (errD_fake + uncond_fake) / 2.
    else:
        errD = errD_real + (errD_fake + errD_wrong) * 0.5
    # the reference returns .item() floats for the three parts (4 host syncs per step); device scalars here
    return errD, errD_real.detach(), errD_wrong.detach(), errD_fake.detach()


def compute_generator_loss(netD, fake_imgs, real_labels, local_label, transf_matrices, transf_matrices_inv, gpus=None,
                           n_objects=3, clamp_negative=False):
    local_label = local_label.detach()
    cond = _label_cond(local_label, n_objects, clamp_negative)
    fake_features = netD(fake_imgs, local_label, transf_matrices, transf_matrices_inv)
    errD_fake = ops.sigmoid_bce(netD.get_cond_logits(fake_features, cond), real_labels, with_logits=True)
    if netD.get_uncond_logits is not None:
        errD_fake = errD_fake + ops.sigmoid_bce(netD.get_uncond_logits(fake_features), real_labels, with_logits=True)
    return errD_fake


def weights_init(m):
    """multi-mnist/miscc/utils.py:127-137 -- N(0, 0.02) by class name."""
    ops.invalidate_packed(m.parameters(recurse=False))     # .data writes below are invisible to torch's version counter
    classname = m.__class__.__name__
    if classname.find('Conv') != -1:
        m.weight.data.normal_(0.0, 0.02)
    elif classname.find('BatchNorm') != -1:
        m.weight.data.normal_(1.0, 0.02)
        m.bias.data.fill_(0)
    elif classname.find('Linear') != -1:
        m.weight.data.normal_(0.0, 0.02)
        if m.bias is not None:
            m.bias.data.fill_(0.0)


def save_model(netG, netD, optimG, optimD, epoch, model_dir, saveD=False, saveOptim=False, max_to_keep=5):
    """multi-mnist/miscc/utils.py:160-174 (same checkpoint dict and rotation)."""
    import glob
    import os
    checkpoint = {'epoch': epoch, 'netG': netG.state_dict(), 'optimG': optimG.state_dict() if saveOptim else {},
                  'netD': netD.state_dict() if saveD else {}, 'optimD': optimD.state_dict() if saveOptim else {}}
    torch.save(checkpoint, "{}/checkpoint_{:04}.pth".format(model_dir, epoch))
    if max_to_keep is not None and max_to_keep > 0:
        ckpts = sorted(glob.glob(model_dir + "/" + '*.pth'))
        while len(ckpts) > max_to_keep:
            os.remove(ckpts[0])
            ckpts = ckpts[1:]


class Stage1Trainer(object):
    """``GANTrainer`` of the single-stage programs -- libmog edition of ``code/multi-mnist/trainer.py`` and
    ``code/clevr/trainer.py`` (``GANTrainer(output_dir)``, ``load_network_stageI``, ``train(data_loader[, stage])``).  The body of
    the hot loop (multi-mnist/trainer.py:134-157, clevr/trainer.py:130-154) is :meth:`train_step`; one process per GPU with an
    NCCL all-reduce per network; the discriminator's never-used weight gradient of the G step is not computed; no ``.item()``
    host syncs in the step.  TensorBoard summaries / image dumps (trainer.py:159-190) and ``sample`` are outside the hot path."""

    program = None     # 'mnist' | 'clevr' (set by the per-program subclass, with cfg / model / losses)

    def __init__(self, output_dir):
        import os
        from . import parallel
        cfg = self.cfg
        parallel.init_from_env()
        ops.precision_from_cfg(cfg)
        if cfg.TRAIN.FLAG and output_dir:
            self.model_dir = os.path.join(output_dir, 'Model')
            self.image_dir = os.path.join(output_dir, 'Image')
            self.log_dir = os.path.join(output_dir, 'Log')
            for d in (self.model_dir, self.image_dir, self.log_dir):
                os.makedirs(d, exist_ok=True)
        self.max_epoch = cfg.TRAIN.MAX_EPOCH
        self.snapshot_interval = cfg.TRAIN.SNAPSHOT_INTERVAL
        self.max_objects = self.n_objects
        self.gpus = [int(ix) for ix in str(cfg.GPU_ID).split(',')]
        self.num_gpus = len(self.gpus)
        self.batch_size = cfg.TRAIN.BATCH_SIZE

    def load_network_stageI(self):
        """multi-mnist/trainer.py:48-72"""
        from . import parallel
        cfg = self.cfg
        netG, netD = self.model.STAGE1_G(), self.model.STAGE1_D()
        netG.apply(weights_init)
        netD.apply(weights_init)
        if cfg.NET_G != '':
            netG.load_state_dict(torch.load(cfg.NET_G, map_location='cpu')["netG"])
        if cfg.NET_D != '':
            netD.load_state_dict(torch.load(cfg.NET_D, map_location='cpu'))
        if cfg.CUDA:
            netG.cuda()
            netD.cuda()
        netG.train()
        netD.train()
        parallel.broadcast_params(netG)
        parallel.broadcast_params(netD)
        return netG, netD

    def define_optimizers(self, netG, netD):
        from . import optim as mog_optim
        cfg = self.cfg
        A = mog_optim.Adam if next(netD.parameters()).is_cuda else torch.optim.Adam
        optimizerD = A(netD.parameters(), lr=cfg.TRAIN.DISCRIMINATOR_LR, betas=(0.5, 0.999))
        optimizerG = A([p for p in netG.parameters() if p.requires_grad], lr=cfg.TRAIN.GENERATOR_LR, betas=(0.5, 0.999))
        return optimizerG, optimizerD

    def make_step_state(self, netG, netD, optimizerG, optimizerD, batch_size=None):
        from . import parallel
        B = batch_size or self.batch_size
        dev = next(netD.parameters()).device
        st = {"netG": netG, "netD": netD, "optG": optimizerG, "optD": optimizerD,
              "real_labels": torch.ones(B, device=dev), "fake_labels": torch.zeros(B, device=dev)}
        if parallel.world() > 1:
            st["bucketG"] = parallel.GradBucket(netG.parameters())
            st["bucketD"] = parallel.GradBucket(netD.parameters())
        return st

    def train_step(self, st, real_imgs, label_one_hot, transf_matrices, transf_matrices_inv, noise=None, optimize=True):
        """One iteration of multi-mnist/trainer.py:134-157 / clevr/trainer.py:130-154.  Returns (errD, errG) device scalars."""
        from . import optim as mog_optim
        from . import parallel
        netG, netD = st["netG"], st["netD"]
        multi = parallel.world() > 1
        B = real_imgs.shape[0]
        if noise is None:
            noise = torch.empty(B, self.cfg.Z_DIM, device=real_imgs.device).normal_(0, 1)
        out = netG(noise, transf_matrices_inv, label_one_hot)
        fake_imgs = out[1] if isinstance(out, tuple) else out
        netD.zero_grad(set_to_none=True)
        errD, _, _, _ = self.losses.compute_discriminator_loss(netD, real_imgs, fake_imgs, st["real_labels"], st["fake_labels"],
                                                               label_one_hot, transf_matrices, transf_matrices_inv, self.gpus)
        errD.backward()      # (the reference passes retain_graph=True; nothing of this graph is used again)
        fusedD = isinstance(st["optD"], mog_optim.Adam)
        if multi:
            st["bucketD"].launch()
            st["bucketD"].finish(scale=not fusedD)
        if optimize:
            if fusedD:
                st["optD"].step(grad_scale=1.0 / parallel.world() if multi else 1.0)
            else:
                st["optD"].step()
        for p in netD.parameters():
            p.requires_grad_(False)
        netG.zero_grad(set_to_none=True)
        errG = self.losses.compute_generator_loss(netD, fake_imgs, st["real_labels"], label_one_hot, transf_matrices,
                                                  transf_matrices_inv, self.gpus)
        errG.backward()
        for p in netD.parameters():
            p.requires_grad_(True)
        fusedG = isinstance(st["optG"], mog_optim.Adam)
        if multi:
            st["bucketG"].launch()
            st["bucketG"].finish(scale=not fusedG)
        if optimize:
            if fusedG:
                st["optG"].step(grad_scale=1.0 / parallel.world() if multi else 1.0)
            else:
                st["optG"].step()
        return errD.detach(), errG.detach()

    def unpack_batch(self, data, dev):
        """DataLoader item -> (real_imgs, label_one_hot, theta, theta^-1) on the device (program specific)."""
        raise NotImplementedError

    def train(self, data_loader, stage=1, max_steps=None):
        import time
        from . import parallel
        cfg = self.cfg
        netG, netD = self.load_network_stageI()
        dev = next(netD.parameters()).device
        optimizerG, optimizerD = self.define_optimizers(netG, netD)
        st = self.make_step_state(netG, netD, optimizerG, optimizerD)
        generator_lr, discriminator_lr = cfg.TRAIN.GENERATOR_LR, cfg.TRAIN.DISCRIMINATOR_LR
        count, epoch = 0, 0
        errD = errG = torch.zeros(())
        for epoch in range(self.max_epoch):
            start_t = time.time()
            if epoch % cfg.TRAIN.LR_DECAY_EPOCH == 0 and epoch > 0:      # trainer.py:107-113
                generator_lr *= 0.5
                discriminator_lr *= 0.5
                for g in optimizerG.param_groups:
                    g['lr'] = generator_lr
                for g in optimizerD.param_groups:
                    g['lr'] = discriminator_lr
            for data in data_loader:
                real_imgs, label_one_hot, tm, tmi = self.unpack_batch(data, dev)
                errD, errG = self.train_step(st, real_imgs, label_one_hot, tm, tmi)
                count += 1
                if max_steps is not None and count >= max_steps:
                    break
            if parallel.rank() == 0:
                print('[%d/%d] Loss_D: %.4f Loss_G: %.4f Total Time: %.2fsec'
                      % (epoch, self.max_epoch, float(errD), float(errG), time.time() - start_t))
                if epoch % self.snapshot_interval == 0 and getattr(self, "model_dir", None):
                    save_model(netG, netD, optimizerG, optimizerD, epoch, self.model_dir)
            if max_steps is not None and count >= max_steps:
                break
        if parallel.rank() == 0 and getattr(self, "model_dir", None):
            save_model(netG, netD, optimizerG, optimizerD, epoch, self.model_dir)
        return st
