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

``build_models`` loads the frozen DAMSM encoders from ``cfg.TRAIN.NET_E`` like trainer.py:56-88 (or takes pre-built
ones); ``train()`` runs as the reference's ``main.py:152`` calls it (no arguments).  ``sampling`` / ``sample`` /
``gen_example`` (trainer.py:387-667) generate images from a checkpoint with the generator in eval mode; the attention-map
collages of ``save_img_results`` / ``build_super_images`` (host-side PIL drawing) are not reproduced.
"""
from __future__ import annotations

import contextlib
import glob
import os
import time

import torch
import torch.optim as optim

from .. import optim as mog_optim
from .. import parallel
from .miscc.config import cfg
from .miscc.losses import KL_loss, class_mask, damsm_terms, discriminator_loss, format_logs, generator_loss
from .miscc.utils import copy_G_params, load_params, mkdir_p, weights_init
from .model import CNN_ENCODER, D_NET64, D_NET128, D_NET256, G_NET, RNN_ENCODER


class condGANTrainer(object):
    def __init__(self, output_dir, data_loader, n_words, ixtoword, resume=False):
        # one process per GPU: under torchrun this joins the process group and selects cuda:LOCAL_RANK (a no-op for a
        # plain ``python main.py`` launch); cfg.MOG.PRECISION selects the conv operand precision (default bf16x3)
        from .. import ops
        parallel.init_from_env()
        ops.precision_from_cfg(cfg)
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
        if cfg.CUDA and torch.cuda.is_available() and parallel.world() == 1:
            torch.cuda.set_device(self.gpus[0])       # trainer.py:50 (torchrun ranks use LOCAL_RANK instead)
        self.n_words = n_words
        self.ixtoword = ixtoword
        self.data_loader = data_loader
        self.num_batches = len(data_loader) if data_loader is not None else 0
        self.device = torch.device('cuda', torch.cuda.current_device()) if torch.cuda.is_available() else None
        self.text_encoder = None
        self.image_encoder = None

    # ------------------------------------------------------------------ models / optimisers
    def _load_encoders(self):
        """trainer.py:56-88 -- frozen CNN_ENCODER / RNN_ENCODER from ``cfg.TRAIN.NET_E`` (the text-encoder checkpoint; the
        image encoder's path is derived from it), eval mode."""
        if cfg.TRAIN.NET_E == '':
            raise RuntimeError("cfg.TRAIN.NET_E is empty: no pretrained DAMSM text / image encoders (trainer.py:56-58). "
                               "Pass encoders to build_models(text_encoder=..., image_encoder=...) or set "
                               "train(damsm=False) explicitly to train without the DAMSM terms.")
        image_encoder = CNN_ENCODER(cfg.TEXT.EMBEDDING_DIM)
        img_encoder_path = cfg.TRAIN.NET_E.replace('text_encoder', 'image_encoder')
        image_encoder.load_state_dict(torch.load(img_encoder_path, map_location='cpu'))
        text_encoder = RNN_ENCODER(self.n_words, nhidden=cfg.TEXT.EMBEDDING_DIM)
        text_encoder.load_state_dict(torch.load(cfg.TRAIN.NET_E, map_location='cpu'))
        for enc in (image_encoder, text_encoder):
            for p in enc.parameters():
                p.requires_grad = False
            enc.eval()
            if cfg.CUDA:
                enc.cuda()
        return text_encoder, image_encoder

    def build_models(self, text_encoder=None, image_encoder=None, load_encoders=True):
        """trainer.py:53-137.  Returns [text_encoder, image_encoder, netG, netsD, epoch].  Encoders that were passed in (here
        or to an earlier call) are kept; otherwise they are loaded from ``cfg.TRAIN.NET_E`` as in the reference."""
        text_encoder = text_encoder if text_encoder is not None else self.text_encoder
        image_encoder = image_encoder if image_encoder is not None else self.image_encoder
        if load_encoders and (text_encoder is None or image_encoder is None):
            te, ie = self._load_encoders()
            text_encoder = text_encoder if text_encoder is not None else te
            image_encoder = image_encoder if image_encoder is not None else ie
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

    def _branch_streams(self, n, device):
        """Side streams for the independent branches of a step (``cfg.MOG.STREAMS``; None = everything on one stream)."""
        if not cfg.MOG.STREAMS or device.type != "cuda":
            return None
        have = getattr(self, "_side_streams", None)
        if have is None or len(have) < n:
            # (MOG_STREAM_PRIO, default on: the branches of many small kernels -- stream 0 = D_NET64, the last = image encoder --
            # at a higher priority than the stream of the big tensor-core kernels)
            prio = os.environ.get("MOG_STREAM_PRIO", "1") == "1"     # (measured: 55.3 -> 54.6 ms)
            lvl = int(os.environ.get("MOG_STREAM_PRIO_LEVELS", "1"))      # (2: big branches above the weight-gradient side streams too)
            have = self._side_streams = [torch.cuda.Stream(device=device, priority=(-lvl if i in (0, n - 1) else -(lvl - 1)) if prio else 0)
                                         for i in range(n)]
        return have

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

    def train_step(self, *args, **kwargs):
        """One iteration of trainer.py:294-342 (see :meth:`_train_step`); with ``cfg.MOG.STREAMS`` the weight gradients run on
        side streams (``ops.async_wgrad``)."""
        from .. import ops
        with ops.async_wgrad(bool(cfg.MOG.STREAMS)):
            return self._train_step(*args, **kwargs)

    def _train_step(self, st, imgs, sent_emb, words_embs, mask, transf_matrices, transf_matrices_inv,
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
        if not multi and os.environ.get("MOG_D_ORDER", "asc") == "desc":     # (tuning knob: enqueue the largest discriminator first)
            order = list(reversed(range(len(netsD))))
        # cfg.MOG.STREAMS: the (independent) discriminator steps run on separate CUDA streams, forked from and joined to the
        # current one (inside a captured step: parallel branches of the graph) -- the small, bandwidth- or latency-bound
        # kernels of one discriminator run next to the tensor-core kernels of another.  Same kernels and sums: bit-identical.
        streams = self._branch_streams(len(netsD) + 1, sent_emb.device)
        cur = torch.cuda.current_stream() if streams is not None else None
        # the DAMSM terms of the generator objective need the generated image and the frozen encoders only: with streams their
        # forward (Inception-v3: ~190 small launches) starts now and runs next to the discriminator steps
        damsm = None
        if streams is not None and self.image_encoder is not None:
            streams[len(netsD)].wait_stream(cur)
            with torch.cuda.stream(streams[len(netsD)]):
                damsm = damsm_terms(self.image_encoder, fake_imgs[len(netsD) - 1], words_embs, sent_emb, st["match_labels"],
                                    cap_lens, class_ids, B)
        errDs = [None] * len(netsD)
        for i in order:
            netD = netsD[i]
            if streams is not None:
                streams[i].wait_stream(cur)
            with (torch.cuda.stream(streams[i]) if streams is not None else contextlib.nullcontext()):
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
                errDs[i] = errD.detach()
        if streams is not None:
            for i in order:
                cur.wait_stream(streams[i])
        for i in range(len(netsD)):
            errD_total = errD_total + errDs[i]
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
                                          transf_matrices_inv=transf_matrices_inv, streams=streams, damsm=damsm)
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

    # ------------------------------------------------------------------ CUDA-graph form of the step
    def graphed_step(self, st, imgs, sent_emb, words_embs, mask, transf_matrices, transf_matrices_inv, label_one_hot,
                     cap_lens=None, class_ids=None, warmup=2, pool=None, noise=None, eps=None, dry_warmup=False):
        """Capture :meth:`train_step` (for these tensor shapes) into one CUDA graph: ~2000 kernel launches, the gradient
        all-reduces and the four fused optimiser steps replay with a single ``cudaGraphLaunch`` and no Python in between
        (the host needs ~40 ms to enqueue a step it takes the device less than that to run).  Returns a
        :class:`GraphedStep`; call it with the next batch's tensors."""
        return GraphedStep(self, st, imgs, sent_emb, words_embs, mask, transf_matrices, transf_matrices_inv, label_one_hot,
                           cap_lens, class_ids, warmup=warmup, pool=pool, noise=noise, eps=eps, dry_warmup=dry_warmup)

    # ------------------------------------------------------------------ epoch loop
    def encode_text(self, text_encoder, captions, cap_lens):
        """trainer.py:281-289 -- frozen text encoder, caption mask trimmed to the longest caption of the batch."""
        hidden = text_encoder.init_hidden(captions.size(0))
        with torch.no_grad():
            words_embs, sent_emb = text_encoder(captions, cap_lens, hidden)
        words_embs, sent_emb = words_embs.detach(), sent_emb.detach()
        mask = (captions == 0)
        num_words = words_embs.size(2)
        if mask.size(1) > num_words:
            mask = mask[:, :num_words]
        return words_embs, sent_emb, mask

    def train(self, prepare_data=None, text_encoder=None, image_encoder=None, damsm=True, max_steps=None):
        """trainer.py:249-366, callable exactly as ``main.py:152`` does (``algo.train()``): the frozen DAMSM encoders come
        from ``cfg.TRAIN.NET_E``, batches go through ``datasets.prepare_data``.  Optional arguments: a custom
        ``prepare_data`` callable, pre-built encoders, ``damsm=False`` to train WITHOUT the DAMSM words / sentence terms
        (never silently: without encoders and without this flag ``build_models`` raises), ``max_steps`` to stop early."""
        if prepare_data is None:
            from .datasets import prepare_data
        text_encoder, image_encoder, netG, netsD, start_epoch = self.build_models(text_encoder, image_encoder)
        if not damsm:
            self.image_encoder = None
        optimizerG, optimizersD = self.define_optimizers(netG, netsD)
        st = self.make_step_state(netG, netsD, optimizerG, optimizersD)
        gen_iterations = 0
        errD = errG = torch.zeros(())
        epoch = start_epoch
        done = False
        use_graph = bool(cfg.MOG.CUDA_GRAPH) and cfg.CUDA and torch.cuda.is_available()
        graphs, pool = {}, None     # one captured graph per caption length T (words_embs is B x nef x T), one memory pool
        for epoch in range(start_epoch, self.max_epoch):
            start_t = time.time()
            for data in self.data_loader:
                imgs, captions, cap_lens, class_ids, keys, tms, label_one_hot = prepare_data(data)
                words_embs, sent_emb, mask = self.encode_text(text_encoder, captions, cap_lens)
                if use_graph:
                    key = (tuple(words_embs.shape), tuple(imgs[-1].shape))
                    g = graphs.get(key)
                    if g is None:
                        if pool is None:
                            pool = torch.cuda.graph_pool_handle()
                        g = graphs[key] = self.graphed_step(st, imgs, sent_emb, words_embs, mask, tms[0], tms[1], label_one_hot,
                                                            cap_lens, class_ids, pool=pool)
                    errD, errG, _ = g(imgs, sent_emb, words_embs, mask, tms[0], tms[1], label_one_hot, cap_lens, class_ids)
                else:
                    errD, errG, _ = self.train_step(st, imgs, sent_emb, words_embs, mask, tms[0], tms[1],
                                                    label_one_hot, cap_lens, class_ids)
                gen_iterations += 1
                if gen_iterations % 1000 == 0 and parallel.rank() == 0:
                    print(format_logs(st["last_logs"]))
                if max_steps is not None and gen_iterations >= max_steps:
                    done = True
                    break
            if parallel.rank() == 0:
                print('[%d/%d][%d] Loss_D: %.2f Loss_G: %.2f Time: %.2fs'
                      % (epoch, self.max_epoch, self.num_batches, float(errD), float(errG), time.time() - start_t))
            if done:
                break
            if epoch % cfg.TRAIN.SNAPSHOT_INTERVAL == 0:
                self.save_model(netG, st["avg_param_G"], netsD, optimizerG, optimizersD, epoch)
        if getattr(self, "model_dir", None):
            self.save_model(netG, st["avg_param_G"], netsD, optimizerG, optimizersD, epoch)
        return st

    # ------------------------------------------------------------------ sampling (generator in eval mode)
    def _load_text_encoder(self):
        text_encoder = RNN_ENCODER(self.n_words, nhidden=cfg.TEXT.EMBEDDING_DIM)
        text_encoder.load_state_dict(torch.load(cfg.TRAIN.NET_E, map_location='cpu'))
        if cfg.CUDA:
            text_encoder.cuda()
        return text_encoder.eval()

    def _load_generator(self):
        """trainer.py:392-420 -- G_NET from the checkpoint dict written by ``save_model`` (``state_dict["netG"]`` = EMA weights)."""
        netG = G_NET()
        state_dict = torch.load(cfg.TRAIN.NET_G, map_location='cpu')
        netG.load_state_dict(state_dict["netG"] if "netG" in state_dict else state_dict)
        if cfg.CUDA:
            netG.cuda()
        return netG.eval()

    @staticmethod
    def _to_uint8_hwc(img_chw):
        """[-1, 1] float CHW -> uint8 HWC exactly as trainer.py:453-457."""
        import numpy as np
        im = img_chw.detach().float().cpu().numpy()
        im = ((im + 1.0) * 127.5).astype(np.uint8)
        return np.transpose(im, (1, 2, 0))

    def save_singleimages(self, images, filenames, save_dir, split_dir, sentenceID=0):
        """trainer.py:368-385"""
        from PIL import Image
        for i in range(images.size(0)):
            s_tmp = '%s/single_samples/%s/%s' % (save_dir, split_dir, filenames[i])
            folder = s_tmp[:s_tmp.rfind('/')]
            if not os.path.isdir(folder):
                mkdir_p(folder)
            img = images[i].add(1).div(2).mul(255).clamp(0, 255).byte()
            Image.fromarray(img.permute(1, 2, 0).contiguous().cpu().numpy()).save('%s_%d.jpg' % (s_tmp, sentenceID))

    def sampling(self, split_dir, num_samples=30000, prepare_data=None):
        """trainer.py:387-459 -- one 256^2 image per caption of the loader, saved as ``<NET_G>/<split>/single/<key>_s-1.png``."""
        from PIL import Image
        if cfg.TRAIN.NET_G == '':
            print('Error: the path for morels is not found!')
            return None
        if prepare_data is None:
            from .datasets import prepare_data
        if split_dir == 'test':
            split_dir = 'valid'
        netG = self._load_generator()
        text_encoder = self.text_encoder if self.text_encoder is not None else self._load_text_encoder()
        nz = cfg.GAN.Z_DIM
        model_dir = cfg.TRAIN.NET_G
        save_dir = '%s/%s' % (model_dir[:model_dir.rfind('.pth')], split_dir)
        mkdir_p(save_dir)
        written = []
        for step, data in enumerate(self.data_loader, 0):
            if step >= num_samples:
                break
            imgs, captions, cap_lens, class_ids, keys, tms, label_one_hot = prepare_data(data)
            words_embs, sent_emb, mask = self.encode_text(text_encoder, captions, cap_lens)
            noise = torch.empty(captions.size(0), nz, device=sent_emb.device).normal_(0, 1)
            with torch.no_grad():
                fake_imgs, _, _, _ = netG(noise, sent_emb, words_embs, mask, tms[1], label_one_hot)
            for j in range(captions.size(0)):
                s_tmp = '%s/single/%s' % (save_dir, keys[j])
                folder = s_tmp[:s_tmp.rfind('/')]
                if not os.path.isdir(folder):
                    mkdir_p(folder)
                k = -1
                fullpath = '%s_s%d.png' % (s_tmp, k)
                Image.fromarray(self._to_uint8_hwc(fake_imgs[k][j])).save(fullpath)
                written.append(fullpath)
        return written

    def sample(self, split_dir, num_samples=25, draw_bbox=False, prepare_data=None):
        """trainer.py:461-579 -- per loader batch: the first caption, 9 noise draws, one row [real | 9 fakes] (optionally with
        the bounding boxes drawn), saved as ``<NET_G>_<split>/<caption>_<step>.png``."""
        import torchvision.utils as vutils
        if cfg.TRAIN.NET_G == '':
            print('Error: the path for model NET_G is not found!')
            return None
        if prepare_data is None:
            from .datasets import prepare_data
        if split_dir == 'test':
            split_dir = 'valid'
        text_encoder = self.text_encoder if self.text_encoder is not None else self._load_text_encoder()
        netG = self._load_generator()
        nz = cfg.GAN.Z_DIM
        model_dir = cfg.TRAIN.NET_G
        save_dir = '%s_%s' % (model_dir[:model_dir.rfind('.pth')], split_dir)
        mkdir_p(save_dir)
        imsize = cfg.TREE.BASE_SIZE * (2 ** (cfg.TREE.BRANCH_NUM - 1))
        written = []
        for step, data in enumerate(self.data_loader, 0):
            if step >= num_samples:
                break
            imgs, captions, cap_lens, class_ids, keys, tms, label_one_hot, bbox = prepare_data(data, eval=True)
            transf_matrices_inv = tms[1][0].unsqueeze(0).repeat(9, 1, 1, 1)
            label9 = label_one_hot[0].unsqueeze(0).repeat(9, 1, 1)
            val_image = imgs[-1][0].reshape(1, 3, imsize, imsize)
            words_embs, sent_emb, mask = self.encode_text(text_encoder, captions, cap_lens)
            words_embs = words_embs[0].unsqueeze(0).repeat(9, 1, 1)
            sent_emb = sent_emb[0].unsqueeze(0).repeat(9, 1)
            mask = mask[0].unsqueeze(0).repeat(9, 1)
            noise = torch.empty(9, nz, device=sent_emb.device).normal_(0, 1)
            with torch.no_grad():
                fake_imgs, _, _, _ = netG(noise, sent_emb, words_embs, mask, transf_matrices_inv, label9)
            data_img = torch.zeros(10, 3, imsize, imsize)
            data_img[0] = val_image.cpu()
            data_img[1:10] = fake_imgs[-1].detach().float().cpu()
            if draw_bbox:
                for idx in range(bbox.shape[1]):
                    x, y, w, h = tuple([int(imsize * float(v)) for v in bbox[0, idx]])
                    w = imsize - 1 if w > imsize - 1 else w
                    h = imsize - 1 if h > imsize - 1 else h
                    if x <= -1:
                        break
                    x2, y2 = min(x + w, imsize - 1), min(y + h, imsize - 1)
                    data_img[:10, :, y, x:x + w] = 1
                    data_img[:10, :, y:y + h, x] = 1
                    data_img[:10, :, y2, x:x + w] = 1
                    data_img[:10, :, y:y + h, x2] = 1
            cap = captions[0].detach().cpu().numpy()
            words = []
            for j in range(len(cap)):
                if cap[j] == 0:
                    break
                words.append(str(self.ixtoword[int(cap[j])]).encode('ascii', 'ignore').decode('ascii'))
            path = '{}/{}_{}.png'.format(save_dir, " ".join(words), step)
            vutils.save_image(data_img, path, normalize=True, nrow=10)
            written.append(path)
        print("Saved {} files to {}".format(len(written), save_dir))
        return written

    def gen_example(self, data_dic, transf_matrices_inv=None, label_one_hot=None):
        """trainer.py:581-667 -- images for hand-written captions: ``data_dic[key] = [cap_array, cap_lens, sorted_indices]``.
        The reference calls ``netG(noise, sent_emb, words_embs, mask)`` here (trainer.py:636), which its own
        ``G_NET.forward`` (six arguments) rejects; this version takes the layout as optional arguments and defaults to an
        EMPTY layout (three empty slots: bbox -1 => theta^-1 = [[-1,0,-4],[0,-1,-4]], label 80), i.e. the global pathway only."""
        from PIL import Image
        from .miscc.utils import compute_transformation_matrix_inverse
        if cfg.TRAIN.NET_G == '':
            print('Error: the path for morels is not found!')
            return None
        text_encoder = self.text_encoder if self.text_encoder is not None else self._load_text_encoder()
        netG = self._load_generator()
        s_tmp = cfg.TRAIN.NET_G[:cfg.TRAIN.NET_G.rfind('.pth')]
        dev = next(netG.parameters()).device
        written = []
        for key in data_dic:
            save_dir = '%s/%s' % (s_tmp, key)
            mkdir_p(save_dir)
            captions, cap_lens, sorted_indices = data_dic[key]
            batch_size = captions.shape[0]
            captions = torch.as_tensor(captions).to(dev)
            cap_lens = torch.as_tensor(cap_lens.copy() if hasattr(cap_lens, "copy") else cap_lens).to(dev)
            tmi, onehot = transf_matrices_inv, label_one_hot
            if tmi is None:
                empty = -torch.ones(batch_size * 3, 4)
                tmi = compute_transformation_matrix_inverse(empty).view(batch_size, 3, 2, 3).to(dev)
            if onehot is None:
                onehot = torch.zeros(batch_size, 3, 81, device=dev)
                onehot[:, :, 80] = 1.0
            words_embs, sent_emb, mask = self.encode_text(text_encoder, captions, cap_lens)
            noise = torch.empty(batch_size, cfg.GAN.Z_DIM, device=dev).normal_(0, 1)
            with torch.no_grad():
                fake_imgs, attention_maps, _, _ = netG(noise, sent_emb, words_embs, mask, tmi, onehot)
            for j in range(batch_size):
                save_name = '%s/%d_s_%d' % (save_dir, 0, sorted_indices[j])
                for k in range(len(fake_imgs)):
                    fullpath = '%s_g%d.png' % (save_name, k)
                    Image.fromarray(self._to_uint8_hwc(fake_imgs[k][j])).save(fullpath)
                    written.append(fullpath)
        return written


class GraphedStep:
    """One ``condGANTrainer.train_step`` captured as a CUDA graph (fixed shapes).  The inputs live in static device buffers
    that every call refreshes (device-to-device or pinned-host-to-device copies on the current stream), the three returned
    losses and ``st["last_logs"]`` are static outputs.  Everything that changes from step to step is device-resident: the
    noise and CA_NET draws come from torch's graph-safe generator, Adam's step count from ``mog_adam_multi_dev``.
    The warm-up steps it runs eagerly before capturing ARE training steps (they update the networks with the given batch)
    unless ``dry_warmup=True``: then they run without the optimiser steps and the BatchNorm running statistics are restored,
    so capturing leaves the training state untouched (needs optimiser state from at least one earlier real step)."""

    def __init__(self, trainer, st, imgs, sent_emb, words_embs, mask, transf_matrices, transf_matrices_inv, label_one_hot,
                 cap_lens=None, class_ids=None, warmup=2, pool=None, noise=None, eps=None, dry_warmup=False):
        self.trainer, self.st = trainer, st
        dev = sent_emb.device
        B = sent_emb.shape[0]
        self.B = B
        self.static = {
            "imgs": [t.detach().to(dev).clone() for t in imgs],
            "sent_emb": sent_emb.detach().clone(), "words_embs": words_embs.detach().clone(), "mask": mask.detach().clone(),
            "transf_matrices": transf_matrices.detach().clone(), "transf_matrices_inv": transf_matrices_inv.detach().clone(),
            "label_one_hot": label_one_hot.detach().clone(),
            "cap_lens": None if cap_lens is None else torch.as_tensor(cap_lens).to(device=dev, dtype=torch.int32).clone(),
            "class_mask": None if class_ids is None else class_mask(class_ids, B, dev).clone(),
            # optional: injected noise / CA_NET draw as static inputs (parity tests); None = drawn inside the graph
            "noise": None if noise is None else noise.detach().clone(), "eps": None if eps is None else eps.detach().clone(),
        }
        s = self.static
        self.opts = [o for o in [st["optG"]] + list(st["optDs"]) if isinstance(o, mog_optim.Adam)]

        def run(optimize=True):
            return trainer.train_step(st, s["imgs"], s["sent_emb"], s["words_embs"], s["mask"], s["transf_matrices"],
                                      s["transf_matrices_inv"], s["label_one_hot"], s["cap_lens"], s["class_mask"],
                                      noise=s["noise"], eps=s["eps"], optimize=optimize)

        self.launches = 0
        saved = None
        if dry_warmup:
            saved = [(bf, bf.detach().clone()) for net in [st["netG"]] + list(st["netsD"]) for bf in net.buffers()]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):     # >= 1: optimiser state and every lazily packed operand must exist
                run(optimize=not dry_warmup)
            if saved is not None:
                for bf, v in saved:
                    bf.copy_(v)
        torch.cuda.current_stream().wait_stream(side)
        # the multi-tensor repacking tables of the optimisers (ops.repack) are (re)built on the host whenever a weight gained a
        # packed layout since the last optimiser step -- e.g. the data-gradient layouts of the discriminators, first needed by
        # the generator step AFTER their own Adam step.  Build them now, eagerly: a rebuild inside the capture would be a
        # pageable host-to-device copy (an error under capture).  Repacking the current weights changes nothing.
        from .. import ops as _ops
        for o in self.opts:
            _ops.repack([p for g in o.param_groups for p in g["params"] if p.grad is not None])
        torch.cuda.synchronize()
        from .. import _lib
        n0 = _lib.launch_count()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, pool=pool):
            self.out = run()
        self.launches = _lib.launch_count() - n0      # libmog kernels per replay
        self.logs = st.get("last_logs")
        # the capture pass executed no kernel, but the host-side step counts moved: take that step back
        for o in self.opts:
            o.advance_host_steps(-1)

    def __call__(self, imgs, sent_emb, words_embs, mask, transf_matrices, transf_matrices_inv, label_one_hot, cap_lens=None,
                 class_ids=None, noise=None, eps=None):
        s = self.static
        for k, v in (("noise", noise), ("eps", eps)):
            if s[k] is not None:
                s[k].copy_(v, non_blocking=True)
        for dst, src in zip(s["imgs"], imgs):
            dst.copy_(src, non_blocking=True)
        for k, v in (("sent_emb", sent_emb), ("words_embs", words_embs), ("mask", mask), ("transf_matrices", transf_matrices),
                     ("transf_matrices_inv", transf_matrices_inv), ("label_one_hot", label_one_hot)):
            s[k].copy_(v, non_blocking=True)
        if s["cap_lens"] is not None:
            s["cap_lens"].copy_(torch.as_tensor(cap_lens), non_blocking=True)
        if s["class_mask"] is not None:
            s["class_mask"].copy_(class_mask(class_ids, self.B, s["class_mask"].device), non_blocking=True)
        self.graph.replay()
        for o in self.opts:
            o.advance_host_steps(1)
        self.st["last_logs"] = self.logs
        return self.out

    def replay(self):
        """Replay on the batch already in the static buffers."""
        self.graph.replay()
        for o in self.opts:
            o.advance_host_steps(1)
        return self.out
