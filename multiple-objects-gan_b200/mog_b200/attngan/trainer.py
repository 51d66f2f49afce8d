"""``condGANTrainer`` -- libmog edition of ``code/coco/attngan/trainer.py`` (training part).

Keeps the reference's class surface (constructor arguments, ``build_models``,
``define_optimizers``, ``prepare_labels``, ``save_model``, ``train``) and checkpoint format
(``{epoch, netG (EMA weights), optimG, netD[], optimD[]}``, trainer.py:173-199).  The body of the
hot loop (trainer.py:294-342) lives in :meth:`train_step`, which bench.py times directly.

Differences, all behind the same results:
* multi-GPU is one process per GPU (``torchrun``/``cfg.GPU_ID``) with one asynchronous NCCL
  all-reduce per network right after its backward (``mog_b200.parallel.GradBucket``) instead of
  single-process ``nn.parallel.data_parallel``;
* discriminator weight gradients are not computed during the G step (the reference computes and
  discards them, trainer.py:271,328);
* no ``.item()`` host syncs inside the step: losses are returned as device scalars.

Sampling / visualisation (trainer.py:201-247, 368-667) and the frozen DAMSM encoders
(``build_models`` 56-84) are outside the hot path; ``build_models`` accepts
pre-built encoders instead of loading them from ``cfg.TRAIN.NET_E``.
"""
from __future__ import annotations

import glob
import os
import time

import torch
import torch.optim as optim

from .. import optim as mog_optim
from .. import parallel
from .miscc.config import cfg
from .miscc.losses import KL_loss, discriminator_loss, format_logs, generator_loss
from .miscc.utils import copy_G_params, load_params, mkdir_p, weights_init
from .model import D_NET64, D_NET128, D_NET256, G_NET


class condGANTrainer(object):
    def __init__(self, output_dir, data_loader, n_words, ixtoword, resume=False):
        if cfg.TRAIN.FLAG and output_dir:
            self.model_dir = os.path.join(output_dir, 'Model')
            self.image_dir = os.path.join(output_dir, 'Image')
            mkdir_p(self.model_dir)
            mkdir_p(self.image_dir)
        self.batch_size = cfg.TRAIN.BATCH_SIZE
        self.max_epoch = cfg.TRAIN.MAX_EPOCH
        self.snapshot_interval = cfg.TRAIN.SNAPSHOT_INTERVAL
        self.resume = resume
        self.gpus = [int(ix) for ix in str(cfg.GPU_ID).split(',')]
        self.n_words = n_words
        self.ixtoword = ixtoword
        self.data_loader = data_loader
        self.num_batches = len(data_loader) if data_loader is not None else 0
        self.device = torch.device('cuda', torch.cuda.current_device()) if torch.cuda.is_available() else None
        self.text_encoder = None
        self.image_encoder = None

    # ------------------------------------------------------------------ models / optimisers
    def build_models(self, text_encoder=None, image_encoder=None):
        """trainer.py:53-137.  Returns [text_encoder, image_encoder, netG, netsD, epoch]."""
        netsD = []
        netG = G_NET()
        if cfg.TREE.BRANCH_NUM > 0:
            netsD.append(D_NET64())
        if cfg.TREE.BRANCH_NUM > 1:
            netsD.append(D_NET128())
        if cfg.TREE.BRANCH_NUM > 2:
            netsD.append(D_NET256())
        if cfg.CUDA:  # initialise on the device: orthogonal_ of D_NET256's 3072x24576 weight is slow on CPU
            netG.cuda()
            for d in netsD:
                d.cuda()
        netG.apply(weights_init)
        for d in netsD:
            d.apply(weights_init)
        epoch = 0
        if self.resume:
            ckpts = sorted(glob.glob(self.model_dir + "/" + '*.pth'))
            latest = ckpts[-1]
            sd = torch.load(latest, map_location='cpu')
            netG.load_state_dict(sd["netG"])
            for i in range(len(netsD)):
                netsD[i].load_state_dict(sd["netD"][i])
            epoch = int(latest[-8:-4]) + 1
        if cfg.TRAIN.NET_G != '':
            netG.load_state_dict(torch.load(cfg.TRAIN.NET_G, map_location='cpu'))
            istart, iend = cfg.TRAIN.NET_G.rfind('_') + 1, cfg.TRAIN.NET_G.rfind('.')
            epoch = int(cfg.TRAIN.NET_G[istart:iend]) + 1
            if cfg.TRAIN.B_NET_D:
                s_tmp = cfg.TRAIN.NET_G[:cfg.TRAIN.NET_G.rfind('/')]
                for i in range(len(netsD)):
                    netsD[i].load_state_dict(torch.load('%s/netD%d.pth' % (s_tmp, i), map_location='cpu'))
        if cfg.CUDA:
            netG.cuda()
            for d in netsD:
                d.cuda()
        netG.train()
        for d in netsD:
            d.train()
            parallel.broadcast_params(d)
        parallel.broadcast_params(netG)
        self.text_encoder, self.image_encoder = text_encoder, image_encoder
        return [text_encoder, image_encoder, netG, netsD, epoch]

    def define_optimizers(self, netG, netsD):
        """trainer.py:139-160 -- Adam(lr, betas=(0.5, 0.999)) per network: the fused libmog optimiser on a GPU
        (same state_dict layout as torch.optim.Adam), torch's own for CPU-resident modules (host-side tests)."""
        on_gpu = next(netG.parameters()).is_cuda
        A = mog_optim.Adam if on_gpu else optim.Adam
        optimizersD = [A(d.parameters(), lr=cfg.TRAIN.DISCRIMINATOR_LR, betas=(0.5, 0.999)) for d in netsD]
        optimizerG = A(netG.parameters(), lr=cfg.TRAIN.GENERATOR_LR, betas=(0.5, 0.999))
        if self.resume:
            ckpts = sorted(glob.glob(self.model_dir + "/" + '*.pth'))
            sd = torch.load(ckpts[-1], map_location='cpu')
            optimizerG.load_state_dict(sd["optimG"])
            for i in range(len(netsD)):
                optimizersD[i].load_state_dict(sd["optimD"][i])
        return optimizerG, optimizersD

    def prepare_labels(self):
        """trainer.py:162-171"""
        dev = self.device if cfg.CUDA else 'cpu'
        real_labels = torch.ones(self.batch_size, device=dev)
        fake_labels = torch.zeros(self.batch_size, device=dev)
        match_labels = torch.arange(self.batch_size, device=dev)
        return real_labels, fake_labels, match_labels

    def save_model(self, netG, avg_param_G, netsD, optimG, optimsD, epoch, max_to_keep=5):
        """trainer.py:173-199 -- checkpoint stores the EMA generator as ``netG``."""
        if parallel.rank() != 0:
            return
        backup_para = copy_G_params(netG)
        load_params(netG, avg_param_G)
        checkpoint = {'epoch': epoch, 'netG': netG.state_dict(), 'optimG': optimG.state_dict(),
                      'netD': [d.state_dict() for d in netsD], 'optimD': [o.state_dict() for o in optimsD]}
        torch.save(checkpoint, "{}/checkpoint_{:04}.pth".format(self.model_dir, epoch))
        load_params(netG, backup_para)
        if max_to_keep is not None and max_to_keep > 0:
            ckpts = sorted(glob.glob(self.model_dir + "/" + '*.pth'))
            while len(ckpts) > max_to_keep:
                os.remove(ckpts[0])
                ckpts = ckpts[1:]

    def set_requires_grad_value(self, models_list, brequires):
        for m in models_list:
            for p in m.parameters():
                p.requires_grad = brequires

    @staticmethod
    def _opt_step(opt, grad_scale=1.0):
        if isinstance(opt, mog_optim.Adam):
            opt.step(grad_scale=grad_scale)
        else:
            opt.step()

    # ------------------------------------------------------------------ the hot loop body
    def make_step_state(self, netG, netsD, optimizerG, optimizersD):
        """Per-run state of :meth:`train_step` (EMA copy, labels, gradient buckets)."""
        st = {"netG": netG, "netsD": netsD, "optG": optimizerG, "optDs": optimizersD,
              "avg_param_G": copy_G_params(netG)}
        st["real_labels"], st["fake_labels"], st["match_labels"] = self.prepare_labels()
        if parallel.world() > 1:
            st["bucketG"] = parallel.GradBucket(netG.parameters())
            st["bucketDs"] = [parallel.GradBucket(d.parameters()) for d in netsD]
        return st

    def train_step(self, st, imgs, sent_emb, words_embs, mask, transf_matrices, transf_matrices_inv,
                   label_one_hot, cap_lens=None, class_ids=None, noise=None, optimize=True, eps=None):
        """One iteration of trainer.py:294-342: G forward; per D: zero_grad, loss, backward, Adam;
        then G: zero_grad, adversarial (+DAMSM if an image encoder is attached) + KL loss, backward,
        Adam, EMA.  Returns (errD_total, errG_total, kl_loss) as device scalars (no host sync)."""
        netG, netsD = st["netG"], st["netsD"]
        multi = parallel.world() > 1
        B = sent_emb.shape[0]
        if noise is None:
            noise = torch.empty(B, cfg.GAN.Z_DIM, device=sent_emb.device).normal_(0, 1)
        fake_imgs, _, mu, logvar = netG(noise, sent_emb, words_embs, mask, transf_matrices_inv, label_one_hot, eps=eps)

        # (3) update the discriminators
        errD_total = 0
        # multi-GPU: the discriminators are independent, so the one with the largest gradient bucket (D_NET256, 643 MB)
        # goes first and its all-reduce travels while the smaller ones compute; single GPU keeps the reference's order
        order = sorted(range(len(netsD)), key=lambda j: -st["bucketDs"][j].flat.numel()) if multi else range(len(netsD))
        for i in order:
            netD = netsD[i]
            netD.zero_grad(set_to_none=True)
            if i == 0:
                errD = discriminator_loss(netD, imgs[i], fake_imgs[i], sent_emb, st["real_labels"],
                                          st["fake_labels"], self.gpus, local_labels=label_one_hot,
                                          transf_matrices=transf_matrices, transf_matrices_inv=transf_matrices_inv)
            else:
                errD = discriminator_loss(netD, imgs[i], fake_imgs[i], sent_emb, st["real_labels"],
                                          st["fake_labels"], self.gpus)
            errD.backward()
            if multi:
                st["bucketDs"][i].launch()       # async all-reduce; next D computes meanwhile
            elif optimize:
                self._opt_step(st["optDs"][i])
            errD_total = errD_total + errD.detach()
        if multi:
            for i in range(len(netsD)):
                fused = isinstance(st["optDs"][i], mog_optim.Adam)
                st["bucketDs"][i].finish(scale=not fused)     # 1/world folded into the fused optimiser pass
                if optimize:
                    self._opt_step(st["optDs"][i], grad_scale=1.0 / parallel.world() if fused else 1.0)

        # (4) update the generator; D weights frozen so their (discarded) wgrad is never computed
        self.set_requires_grad_value(netsD, False)
        netG.zero_grad(set_to_none=True)
        errG_total, logs = generator_loss(netsD, self.image_encoder, fake_imgs, st["real_labels"], words_embs,
                                          sent_emb, st["match_labels"], cap_lens, class_ids, self.gpus,
                                          local_labels=label_one_hot, transf_matrices=transf_matrices,
                                          transf_matrices_inv=transf_matrices_inv)
        kl_loss = KL_loss(mu, logvar)
        errG_total = errG_total + kl_loss
        errG_total.backward()
        self.set_requires_grad_value(netsD, True)
        fusedG = isinstance(st["optG"], mog_optim.Adam)
        if multi:
            st["bucketG"].launch()
            st["bucketG"].finish(scale=not fusedG)
        if optimize:
            if fusedG:   # Adam + EMA (trainer.py:340-342) in one pass over the parameters
                st["optG"].step(ema_params=st["avg_param_G"], ema_decay=0.999,
                                grad_scale=1.0 / parallel.world() if multi else 1.0)
            else:
                st["optG"].step()
                with torch.no_grad():   # EMA, trainer.py:341-342
                    params = list(netG.parameters())
                    torch._foreach_mul_(st["avg_param_G"], 0.999)
                    torch._foreach_add_(st["avg_param_G"], [p.data for p in params], alpha=0.001)
        st["last_logs"] = logs
        return errD_total, errG_total.detach(), kl_loss.detach()

    # ------------------------------------------------------------------ epoch loop
    def train(self, prepare_data=None, text_encoder=None, image_encoder=None):
        """trainer.py:249-366.  ``data_loader`` must yield what the reference's ``prepare_data``
        consumes (datasets.py:28-68) unless a custom ``prepare_data`` callable is given; caption
        embeddings come from ``text_encoder`` (frozen)."""
        text_encoder, image_encoder, netG, netsD, start_epoch = self.build_models(text_encoder, image_encoder)
        optimizerG, optimizersD = self.define_optimizers(netG, netsD)
        st = self.make_step_state(netG, netsD, optimizerG, optimizersD)
        gen_iterations = 0
        errD = errG = torch.zeros(())
        epoch = start_epoch
        for epoch in range(start_epoch, self.max_epoch):
            start_t = time.time()
            for data in self.data_loader:
                imgs, captions, cap_lens, class_ids, keys, tms, label_one_hot = prepare_data(data)
                hidden = text_encoder.init_hidden(self.batch_size)
                with torch.no_grad():
                    words_embs, sent_emb = text_encoder(captions, cap_lens, hidden)
                mask = (captions == 0)
                if mask.size(1) > words_embs.size(2):
                    mask = mask[:, :words_embs.size(2)]
                errD, errG, _ = self.train_step(st, imgs, sent_emb, words_embs, mask, tms[0], tms[1],
                                                label_one_hot, cap_lens, class_ids)
                gen_iterations += 1
                if gen_iterations % 1000 == 0 and parallel.rank() == 0:
                    print(format_logs(st["last_logs"]))
            if parallel.rank() == 0:
                print('[%d/%d][%d] Loss_D: %.2f Loss_G: %.2f Time: %.2fs'
                      % (epoch, self.max_epoch, self.num_batches, float(errD), float(errG), time.time() - start_t))
            if epoch % cfg.TRAIN.SNAPSHOT_INTERVAL == 0:
                self.save_model(netG, st["avg_param_G"], netsD, optimizerG, optimizersD, epoch)
        self.save_model(netG, st["avg_param_G"], netsD, optimizerG, optimizersD, epoch)
