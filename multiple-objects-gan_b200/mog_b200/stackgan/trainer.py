"""``GANTrainer`` -- libmog edition of ``code/coco/stackgan/trainer.py`` (training part).

Keeps the reference's class surface (``GANTrainer(output_dir)``, ``load_network_stageI``,
``load_network_stageII``, ``train(data_loader, stage, max_objects)``) and checkpoint format
(``{epoch, netG, optimG, netD, optimD}``, miscc/utils.py:162-176).  The body of the hot loop
(trainer.py:154-235) lives in :meth:`train_step`.

Differences, all behind the same results: one process per GPU with an NCCL all-reduce per network
(``mog_b200.parallel.GradBucket``) instead of ``nn.parallel.data_parallel``; the discriminator's
weight gradients are not computed during the G step (the reference computes them and never uses
them: ``netD.zero_grad()`` precedes the next D step, trainer.py:204); no ``.item()`` host syncs in
the step.  TensorBoard summaries and image dumps (trainer.py:237-266) are outside the hot path; ``sample`` (:287-420)
generates the reference's [real | 9 fakes] rows from a checkpoint with the generator in eval mode.
"""
from __future__ import annotations

import os
import time

import torch
import torch.optim as optim

from .. import optim as mog_optim
from .. import parallel
from .miscc.config import cfg
from .miscc.utils import (KL_loss, compute_discriminator_loss, compute_generator_loss, compute_transformation_matrix,
                          compute_transformation_matrix_inverse, mkdir_p, save_model, weights_init)


class GANTrainer(object):
    def __init__(self, output_dir):
        from .. import ops
        parallel.init_from_env()          # one process per GPU under torchrun (no-op for a plain launch)
        ops.precision_from_cfg(cfg)
        if cfg.TRAIN.FLAG and output_dir:
            self.model_dir = os.path.join(output_dir, 'Model')
            self.image_dir = os.path.join(output_dir, 'Image')
            self.log_dir = os.path.join(output_dir, 'Log')
            for d in (self.model_dir, self.image_dir, self.log_dir):
                mkdir_p(d)
        self.max_epoch = cfg.TRAIN.MAX_EPOCH
        self.snapshot_interval = cfg.TRAIN.SNAPSHOT_INTERVAL
        self.gpus = [int(ix) for ix in str(cfg.GPU_ID).split(',')]
        self.num_gpus = len(self.gpus)
        self.batch_size = cfg.TRAIN.BATCH_SIZE

    # ------------------------------------------------------------------ networks
    def _finish(self, netG, netD):
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

    def load_network_stageI(self):
        """trainer.py:50-72"""
        from .model import STAGE1_D, STAGE1_G
        netG = STAGE1_G()
        netG.apply(weights_init)
        netD = STAGE1_D()
        netD.apply(weights_init)
        if cfg.NET_G != '':
            netG.load_state_dict(torch.load(cfg.NET_G, map_location='cpu')["netG"])
        return self._finish(netG, netD)

    def load_network_stageII(self):
        """trainer.py:75-108 -- needs a stage-I generator (``cfg.STAGE1_G``) or a full stage-II checkpoint."""
        from .model import STAGE1_G, STAGE2_D, STAGE2_G
        netG = STAGE2_G(STAGE1_G())
        netG.apply(weights_init)
        if cfg.NET_G != '':
            netG.load_state_dict(torch.load(cfg.NET_G, map_location='cpu')["netG"])
        elif cfg.STAGE1_G != '':
            netG.STAGE1_G.load_state_dict(torch.load(cfg.STAGE1_G, map_location='cpu')["netG"])
        else:
            print("Please give the Stage1_G path")
            return
        netD = STAGE2_D()
        netD.apply(weights_init)
        return self._finish(netG, netD)

    def define_optimizers(self, netG, netD):
        """trainer.py:131-137 -- Adam(betas=(0.5, 0.999)); only the trainable generator parameters."""
        A = mog_optim.Adam if next(netD.parameters()).is_cuda else optim.Adam   # fused libmog Adam on a GPU
        optimizerD = A(netD.parameters(), lr=cfg.TRAIN.DISCRIMINATOR_LR, betas=(0.5, 0.999))
        optimizerG = A([p for p in netG.parameters() if p.requires_grad], lr=cfg.TRAIN.GENERATOR_LR, betas=(0.5, 0.999))
        return optimizerG, optimizerD

    # ------------------------------------------------------------------ the hot loop body
    def make_step_state(self, netG, netD, optimizerG, optimizerD, batch_size=None):
        B = batch_size or self.batch_size
        dev = next(netD.parameters()).device
        st = {"netG": netG, "netD": netD, "optG": optimizerG, "optD": optimizerD,
              "real_labels": torch.ones(B, device=dev), "fake_labels": torch.zeros(B, device=dev)}
        if parallel.world() > 1:
            st["bucketG"] = parallel.GradBucket(netG.parameters())
            st["bucketD"] = parallel.GradBucket(netD.parameters())
        return st

    def train_step(self, st, real_imgs, txt_embedding, label_one_hot, transf_matrices_inv, transf_matrices=None,
                   transf_matrices_s2=None, transf_matrices_inv_s2=None, noise=None, optimize=True, stage=None):
        """One iteration of trainer.py:193-235.  Stage I: ``transf_matrices`` / ``transf_matrices_inv``; stage II
        additionally the ``_s2`` pair (the D and the stage-II object pathway use it).  Returns
        (errD, errG, kl_loss) as device scalars."""
        stage = cfg.STAGE if stage is None else stage
        netG, netD = st["netG"], st["netD"]
        multi = parallel.world() > 1
        B = real_imgs.shape[0]
        if noise is None:
            noise = torch.empty(B, cfg.Z_DIM, device=real_imgs.device).normal_(0, 1)
        if stage == 1:
            _, fake_imgs, mu, logvar, _ = netG(txt_embedding, noise, transf_matrices_inv, label_one_hot)
            th, thi = transf_matrices, transf_matrices_inv
        else:
            _, fake_imgs, mu, logvar, _ = netG(txt_embedding, noise, transf_matrices_inv, transf_matrices_s2,
                                               transf_matrices_inv_s2, label_one_hot)
            th, thi = transf_matrices_s2, transf_matrices_inv_s2
        # (3) update D
        netD.zero_grad(set_to_none=True)
        errD, _, _, _ = compute_discriminator_loss(netD, real_imgs, fake_imgs, st["real_labels"], st["fake_labels"],
                                                   label_one_hot, th, thi, mu, self.gpus)
        errD.backward()   # (the reference passes retain_graph=True; nothing of this graph is used again)
        if multi:
            st["bucketD"].launch()
            st["bucketD"].finish()
        if optimize:
            st["optD"].step()
        # (2) update G; D weights frozen so their (never used) wgrad is not computed
        for p in netD.parameters():
            p.requires_grad_(False)
        netG.zero_grad(set_to_none=True)
        errG = compute_generator_loss(netD, fake_imgs, st["real_labels"], label_one_hot, th, thi, mu, self.gpus)
        kl_loss = KL_loss(mu, logvar)
        (errG + kl_loss * cfg.TRAIN.COEFF.KL).backward()
        for p in netD.parameters():
            p.requires_grad_(True)
        if multi:
            st["bucketG"].launch()
            st["bucketG"].finish()
        if optimize:
            st["optG"].step()
        return errD.detach(), errG.detach(), kl_loss.detach()

    # ------------------------------------------------------------------ epoch loop
    def train(self, data_loader, stage=1, max_objects=3):
        """trainer.py:110-283.  ``data_loader`` yields (real_img, bbox, label, txt_embedding) like the reference's
        ``TextDataset`` (stage II: bbox is the [stage-I-scaled, stage-II-scaled] pair)."""
        nets = self.load_network_stageI() if stage == 1 else self.load_network_stageII()
        if nets is None:
            return
        netG, netD = nets
        dev = next(netD.parameters()).device
        optimizerG, optimizerD = self.define_optimizers(netG, netD)
        st = self.make_step_state(netG, netD, optimizerG, optimizerD)
        generator_lr, discriminator_lr = cfg.TRAIN.GENERATOR_LR, cfg.TRAIN.DISCRIMINATOR_LR
        epoch = 0
        for epoch in range(self.max_epoch):
            start_t = time.time()
            if epoch % cfg.TRAIN.LR_DECAY_EPOCH == 0 and epoch > 0:   # trainer.py:141-147
                generator_lr *= 0.5
                discriminator_lr *= 0.5
                for g in optimizerG.param_groups:
                    g['lr'] = generator_lr
                for g in optimizerD.param_groups:
                    g['lr'] = discriminator_lr
            errD = errG = kl = torch.zeros(())
            for real_img_cpu, bbox, label, txt_embedding in data_loader:
                real_imgs = real_img_cpu.to(dev, non_blocking=True)
                txt_embedding = txt_embedding.to(dev, non_blocking=True).float()
                B = real_imgs.shape[0]
                kw = {}
                if stage == 1:
                    bb = bbox.to(dev).view(-1, 4).float()
                    tinv = compute_transformation_matrix_inverse(bb).view(B, max_objects, 2, 3)
                    kw["transf_matrices"] = compute_transformation_matrix(bb).view(B, max_objects, 2, 3)
                else:
                    b1, b2 = bbox[0].to(dev).view(-1, 4).float(), bbox[1].to(dev).view(-1, 4).float()
                    tinv = compute_transformation_matrix_inverse(b1).view(B, max_objects, 2, 3)
                    kw["transf_matrices_inv_s2"] = compute_transformation_matrix_inverse(b2).view(B, max_objects, 2, 3)
                    kw["transf_matrices_s2"] = compute_transformation_matrix(b2).view(B, max_objects, 2, 3)
                _labels = label.to(dev).long().clone()
                _labels[_labels < 0] = 80                                # trainer.py:186-189
                label_one_hot = torch.zeros(B, max_objects, 81, device=dev).scatter_(2, _labels.view(B, max_objects, 1), 1.0)
                errD, errG, kl = self.train_step(st, real_imgs, txt_embedding, label_one_hot, tinv, stage=stage, **kw)
            if parallel.rank() == 0:
                print('[%d/%d] Loss_D: %.4f Loss_G: %.4f Loss_KL: %.4f Total Time: %.2fsec'
                      % (epoch, self.max_epoch, float(errD), float(errG), float(kl), time.time() - start_t))
                if epoch % self.snapshot_interval == 0:
                    save_model(netG, netD, optimizerG, optimizerD, epoch, self.model_dir)
        if parallel.rank() == 0:
            save_model(netG, netD, optimizerG, optimizerD, epoch, self.model_dir)

    # ------------------------------------------------------------------ sampling (generator in eval mode)
    def sample(self, datapath, num_samples=25, stage=1, draw_bbox=True, max_objects=3, data=None):
        """trainer.py:287-420 -- ``num_samples`` rows [validation image | 9 generated images] for random validation captions,
        saved to ``<NET_G>_visualize_bbox/<caption>.png``.  ``data`` (optional) supplies what the reference reads from
        ``datapath`` (``val_captions.t7`` via torchfile, ``filenames.pickle``, the label / bbox pickles, the jpg images):
        ``{"embeddings" [n,1024], "captions" [n], "label" [n,3,1], "bbox" [n,3,4] (stage II: a pair), "images" [n,3,S,S] or None}``."""
        import numpy as np
        import torchvision.utils as vutils
        nets = self.load_network_stageI() if stage == 1 else self.load_network_stageII()
        if nets is None:
            return None
        netG = nets[0].eval()
        dev = next(netG.parameters()).device
        if data is None:
            data = self._load_sample_data(datapath, stage)
        embeddings = np.asarray(data["embeddings"], np.float32)
        captions_list, n = list(data["captions"]), len(data["captions"])
        label = torch.as_tensor(data["label"]).to(dev)
        bbox = data["bbox"]
        bbox = [torch.as_tensor(b).to(dev).float() for b in bbox] if stage == 2 else torch.as_tensor(bbox).to(dev).float()
        bbox_ = (bbox[0] if stage == 2 else bbox).clone()
        first = bbox[0] if stage == 2 else bbox
        tinv = compute_transformation_matrix_inverse(first.view(-1, 4)).view(n, max_objects, 2, 3)
        if stage == 2:
            tinv2 = compute_transformation_matrix_inverse(bbox[1].view(-1, 4)).view(n, max_objects, 2, 3)
            t2 = compute_transformation_matrix(bbox[1].view(-1, 4)).view(n, max_objects, 2, 3)
        _labels = label.long().view(n, max_objects, 1).clone()
        _labels[_labels < 0] = 80
        label_one_hot = torch.zeros(n, max_objects, 81, device=dev).scatter_(2, _labels, 1.0)
        save_dir = cfg.NET_G[:cfg.NET_G.find('.pth')] + "_visualize_bbox"
        mkdir_p(save_dir)
        imsize = 64 if stage == 1 else 256
        written = []
        for count in range(num_samples):
            index = int(np.random.randint(0, n, 1)[0])
            val_image = torch.zeros(1, 3, imsize, imsize)
            if data.get("images") is not None:
                val_image = torch.as_tensor(data["images"][index]).float().view(1, 3, imsize, imsize)
            txt = torch.from_numpy(np.reshape(embeddings[index], (1, -1)).repeat(9, 0)).to(dev)
            rep = lambda t: t[index].view(1, max_objects, 2, 3).repeat(9, 1, 1, 1)   # noqa: E731
            onehot9 = label_one_hot[index].view(1, max_objects, 81).repeat(9, 1, 1)
            noise = torch.empty(9, cfg.Z_DIM, device=dev).normal_(0, 1)
            with torch.no_grad():
                if stage == 1:
                    _, fake_imgs, _, _, _ = netG(txt, noise, rep(tinv), onehot9)
                else:
                    _, fake_imgs, _, _, _ = netG(txt, noise, rep(tinv), rep(t2), rep(tinv2), onehot9)
            data_img = torch.zeros(10, 3, imsize, imsize)
            data_img[0] = val_image
            data_img[1:10] = fake_imgs.detach().float().cpu()
            if draw_bbox:
                for idx in range(max_objects):
                    x, y, w, h = tuple([int(imsize * float(v)) for v in bbox_[index, idx]])
                    w = imsize - 1 if w > imsize - 1 else w
                    h = imsize - 1 if h > imsize - 1 else h
                    if x <= -1:
                        break
                    x2, y2 = min(x + w, imsize - 1), min(y + h, imsize - 1)
                    data_img[:10, :, y, x:x + w] = 1
                    data_img[:10, :, y:y + h, x] = 1
                    data_img[:10, :, y2, x:x + w] = 1
                    data_img[:10, :, y:y + h, x2] = 1
            path = '{}/{}.png'.format(save_dir, str(captions_list[index]).replace("/", "_")[:150])
            vutils.save_image(data_img, path, normalize=True, nrow=10)
            written.append(path)
        print("Saved {} files to {}".format(len(written), save_dir))
        return written

    @staticmethod
    def _load_sample_data(datapath, stage):
        """The validation files of trainer.py:299-309 (needs the ``torchfile`` package for ``val_captions.t7``)."""
        import pickle
        import numpy as np
        try:
            import torchfile
        except ImportError as e:
            raise RuntimeError("sample(): reading %sval_captions.t7 needs the `torchfile` package; pass data=... instead" % datapath) from e
        t_file = torchfile.load(datapath + "val_captions.t7")
        with open(os.path.join(datapath, 'bboxes.pickle'), 'rb') as f:
            bbox = np.asarray(pickle.load(f, encoding='latin1'))
        with open(os.path.join(datapath, 'labels.pickle'), 'rb') as f:
            label = np.asarray(pickle.load(f, encoding='latin1'))
        return {"embeddings": np.concatenate(t_file.fea_txt, axis=0), "captions": list(t_file.raw_txt), "label": label,
                "bbox": [bbox, bbox] if stage == 2 else bbox, "images": None}
