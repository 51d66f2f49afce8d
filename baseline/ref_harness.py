"""Harness that imports and drives the UNMODIFIED reference (tohinz/multiple-objects-gan).

Test / measurement infrastructure only -- nothing under ``multiple-objects-gan_b200/`` imports it.
Users: ``bench.py --impl reference`` and the ``cpu_baseline`` / ``torch_cudnn_b200`` legs of
``bench.py`` (timing), ``tests/golden/make_golden_trainstep.py`` (fixtures).

The reference sources are never committed: ``__graft_entry__.build()`` copies
``/root/reference/code`` into the git-ignored ``baseline/_ref/code`` (it travels to the GPU box with
the repo snapshot, like the built ``.so``); in the build container ``/root/reference/code`` is used
directly when the copy is absent.

Shims -- all outside the reference's arithmetic:
  1-3  ``easydict`` / ``skimage`` / ``cPickle`` / ``nltk`` stand-ins on ``sys.path`` (``baseline/shims``)
  4    CPU only: ``torch.cuda.FloatTensor -> torch.FloatTensor`` (canvases are allocated with the CUDA
       constructor inside ``forward``, model.py:106,388,391,684)
  5    ``torch.ByteTensor(masks)`` -> bool tensor (uint8 masks are rejected by ``masked_fill_`` today)
  6    CPU only: ``nn.parallel.data_parallel`` -> direct call
  7    ``model_zoo.load_url`` -> deterministic stand-in for the ImageNet Inception-v3 weights (no network)
  8    ``cfg`` fields set in code (``cfg_from_file`` is py2-only)
"""
from __future__ import annotations

import importlib
import os
import sys
import time

import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SHIMS = os.path.join(HERE, "shims")

PROGRAMS = {"attngan": "coco/attngan", "stackgan": "coco/stackgan", "mnist": "multi-mnist", "clevr": "clevr"}


def ref_code_dir():
    for base in (os.path.join(HERE, "_ref", "code"), "/root/reference/code"):
        if os.path.isdir(base):
            return base
    return None


def available():
    return ref_code_dir() is not None


_loaded = {}


def load(program="attngan", device="cpu"):
    """Import the reference modules of one program; returns a namespace with cfg, model, losses, utils (+GlobalAttention).
    Only one program can be live per process (they all use the top-level module names ``model`` / ``miscc``)."""
    if _loaded:
        if program in _loaded:
            return _loaded[program]
        raise RuntimeError("the reference programs share module names: one program per process")
    base = ref_code_dir()
    if base is None:
        raise RuntimeError("reference sources not found (baseline/_ref/code or /root/reference/code)")
    sys.path[:0] = [SHIMS, os.path.join(base, PROGRAMS[program]), os.path.join(ROOT, "multiple-objects-gan_b200")]
    cpu = str(device) == "cpu"
    if cpu:
        torch.cuda.FloatTensor = torch.FloatTensor            # shim 4
        torch.cuda.DoubleTensor = torch.DoubleTensor

        def _dp(module, inputs, device_ids=None, **kw):          # shim 6
            return module(*inputs) if isinstance(inputs, tuple) else module(inputs)
        nn.parallel.data_parallel = _dp
    torch.ByteTensor = lambda a: torch.as_tensor(a).bool()      # shim 5

    class NS:
        pass
    ns = NS()
    ns.program, ns.cpu = program, cpu
    ns.cfg = importlib.import_module("miscc.config").cfg
    ns.cfg.CUDA = not cpu
    ns.model = importlib.import_module("model")
    ns.utils = importlib.import_module("miscc.utils")
    if program == "attngan":
        ns.losses = importlib.import_module("miscc.losses")
        ns.GlobalAttention = importlib.import_module("GlobalAttention")
    _loaded[program] = ns
    return ns


# ------------------------------------------------------------------------------------------------
# AttnGAN (config 5 of BASELINE.json): trainer.py:294-342 on synthetic data
# ------------------------------------------------------------------------------------------------
CFG5 = dict(GF_DIM=48, DF_DIM=96, Z_DIM=100, R_NUM=3, EMBEDDING_DIM=256, T=18)


def set_attngan_cfg(ns, c, B):
    cfg = ns.cfg
    cfg.GAN.GF_DIM, cfg.GAN.DF_DIM, cfg.GAN.Z_DIM = c["GF_DIM"], c["DF_DIM"], c["Z_DIM"]
    cfg.GAN.CONDITION_DIM, cfg.GAN.R_NUM = 100, c["R_NUM"]
    cfg.TEXT.EMBEDDING_DIM, cfg.TEXT.WORDS_NUM = c["EMBEDDING_DIM"], c["T"]
    cfg.TREE.BRANCH_NUM = 3
    cfg.TRAIN.BATCH_SIZE = B
    cfg.TRAIN.FLAG = True
    cfg.TRAIN.SMOOTH.GAMMA1, cfg.TRAIN.SMOOTH.GAMMA2 = 4.0, 5.0
    cfg.TRAIN.SMOOTH.GAMMA3, cfg.TRAIN.SMOOTH.LAMBDA = 10.0, 50.0
    cfg.TRAIN.DISCRIMINATOR_LR = cfg.TRAIN.GENERATOR_LR = 2e-4


def reference_cnn_encoder(ns, nef, seed=9):
    """The reference's own CNN_ENCODER (model.py:207-313) over torchvision's Inception-v3, with shim 7."""
    import torch.utils.model_zoo as model_zoo
    import torchvision
    from mog_b200 import synth
    full = torchvision.models.inception_v3(weights=None, aux_logits=True, init_weights=False).state_dict()
    full = synth.fill_encoder_state_dict(full, seed)
    model_zoo.load_url = lambda url, *a, **k: full
    import io
    import contextlib
    with contextlib.redirect_stdout(io.StringIO()):
        enc = ns.model.CNN_ENCODER(nef)
    return enc


class AttnGANStep:
    """State + body of one iteration of ``code/coco/attngan/trainer.py:294-342`` with the reference's own modules,
    losses, ``optim.Adam(betas=(0.5, 0.999))`` and EMA, on the synthetic batch of ``mog_b200.synth``."""

    def __init__(self, ns, B, c=None, seed=1234, device="cpu", damsm=True, encoder="reference", init="weights_init",
                 fill_seed=None, logit_scale=None):
        import torch.optim as optim
        from mog_b200 import synth
        self.ns, self.B, self.device = ns, B, torch.device(device)
        c = dict(CFG5 if c is None else c)
        self.c = c
        set_attngan_cfg(ns, c, B)
        M, U = ns.model, ns.utils
        torch.manual_seed(seed)
        self.netG = M.G_NET()
        self.netsD = [M.D_NET64(), M.D_NET128(), M.D_NET256()]
        if init == "weights_init":       # trainer.py:99-103
            if self.device.type == "cuda":   # orthogonal_ of D_NET256's 3072 x 24576 weight takes minutes on a CPU
                self.netG.to(self.device)
                for d in self.netsD:
                    d.to(self.device)
            self.netG.apply(U.weights_init)
            for d in self.netsD:
                d.apply(U.weights_init)
        elif init == "fill":             # deterministic numpy weights shared with the libmog side of a parity test
            s = seed if fill_seed is None else fill_seed
            self.netG.load_state_dict(synth.fill_state_dict(self.netG.state_dict(), s + 1))
            for i, d in enumerate(self.netsD):
                sd = synth.fill_state_dict(d.state_dict(), s + 2 + i)
                if logit_scale is not None:
                    synth.soften_logits(sd, logit_scale)
                d.load_state_dict(sd)
        elif init == "fast":             # N(0, 1/fan_in): what the CPU timing arm uses (orthogonal init is slow there)
            self.netG.load_state_dict(synth.fill_state_dict(self.netG.state_dict(), 1))
            for i, d in enumerate(self.netsD):
                d.load_state_dict(synth.fill_state_dict(d.state_dict(), 2 + i))
        self.image_encoder = None
        if damsm:
            if encoder == "reference":
                enc = reference_cnn_encoder(ns, c["EMBEDDING_DIM"])
                own = synth.fill_encoder_state_dict({k: v for k, v in enc.state_dict().items() if k.startswith("emb_")}, 10)
                enc.load_state_dict(dict(enc.state_dict(), **own))
                for p in enc.parameters():
                    p.requires_grad = False
                self.image_encoder = enc.to(self.device).eval()       # trainer.py:66-69
            else:
                self.image_encoder = encoder
        self.netG.to(self.device).train()
        for d in self.netsD:
            d.to(self.device).train()
        self.avg_param_G = U.copy_G_params(self.netG)              # trainer.py:251
        self.optimizersD = [optim.Adam(d.parameters(), lr=ns.cfg.TRAIN.DISCRIMINATOR_LR, betas=(0.5, 0.999))
                            for d in self.netsD]                    # trainer.py:141-148
        self.optimizerG = optim.Adam(self.netG.parameters(), lr=ns.cfg.TRAIN.GENERATOR_LR, betas=(0.5, 0.999))
        dev = self.device
        self.real_labels = torch.ones(B, device=dev)               # trainer.py:162-171
        self.fake_labels = torch.zeros(B, device=dev)
        self.match_labels = torch.arange(B, device=dev)
        b = synth.attngan_batch(B, T=c["T"], nef=c["EMBEDDING_DIM"], nz=c["Z_DIM"], seed=seed)
        self.batch = b
        self.d = {k: v.to(dev) for k, v in b.items() if torch.is_tensor(v)}
        self.imgs = [t.to(dev) for t in b["imgs"]]
        self.noise = torch.empty(B, c["Z_DIM"], device=dev)
        self.gpus = [0]

    def step(self, noise=None):
        ns, L = self.ns, self.ns.losses
        d, netG, netsD = self.d, self.netG, self.netsD
        if noise is None:
            self.noise.normal_(0, 1)                                 # trainer.py:294
        else:
            self.noise.copy_(noise)
        tm, tmi, onehot = d["transf_matrices"], d["transf_matrices_inv"], d["label_one_hot"]
        inputs = (self.noise, d["sent_emb"], d["words_embs"], d["mask"], tmi, onehot)
        fake_imgs, _, mu, logvar = nn.parallel.data_parallel(netG, inputs, self.gpus)
        errD_total = 0
        for i in range(len(netsD)):                                  # trainer.py:301-318
            netsD[i].zero_grad()
            if i == 0:
                errD = L.discriminator_loss(netsD[i], self.imgs[i], fake_imgs[i], d["sent_emb"], self.real_labels,
                                            self.fake_labels, self.gpus, local_labels=onehot, transf_matrices=tm,
                                            transf_matrices_inv=tmi)
            else:
                errD = L.discriminator_loss(netsD[i], self.imgs[i], fake_imgs[i], d["sent_emb"], self.real_labels,
                                            self.fake_labels, self.gpus)
            errD.backward()
            self.optimizersD[i].step()
            errD_total = errD_total + errD.detach()
        netG.zero_grad()                                             # trainer.py:329-342
        if self.image_encoder is not None:
            errG_total, _ = L.generator_loss(netsD, self.image_encoder, fake_imgs, self.real_labels, d["words_embs"],
                                             d["sent_emb"], self.match_labels, d["cap_lens"], self.batch["class_ids"],
                                             self.gpus, local_labels=onehot, transf_matrices=tm, transf_matrices_inv=tmi)
        else:   # G+D-only variant: generator_loss with the ranking branch removed (losses.py:189-204)
            errG_total = 0
            bce = nn.BCELoss()
            for i, netD in enumerate(netsD):
                f = nn.parallel.data_parallel(netD, (fake_imgs[i], onehot, tm, tmi) if i == 0 else (fake_imgs[i]), self.gpus)
                errG_total = errG_total + bce(netD.UNCOND_DNET(f), self.real_labels) + \
                    bce(netD.COND_DNET(f, d["sent_emb"]), self.real_labels)
        kl_loss = L.KL_loss(mu, logvar)
        errG_total = errG_total + kl_loss
        errG_total.backward()
        self.optimizerG.step()
        for p, avg_p in zip(netG.parameters(), self.avg_param_G):
            avg_p.mul_(0.999).add_(p.data, alpha=0.001)
        return errD_total, errG_total.detach(), kl_loss.detach(), fake_imgs


def time_attngan(B, steps, warmup, device="cpu", damsm=True, tf32=None, init=None, threads=None):
    """Wall-clock (CPU) / CUDA-event (GPU) time of `steps` iterations after `warmup`; returns a dict."""
    import torch.backends.cudnn as cudnn
    dev = torch.device(device)
    ns = load("attngan", dev.type)
    if dev.type == "cpu":
        threads = threads or (os.cpu_count() or 1)
        torch.set_num_threads(threads)
    else:
        cudnn.benchmark = True                                       # trainer.py:51
        if tf32 is not None:
            torch.backends.cudnn.allow_tf32 = bool(tf32)
            torch.backends.cuda.matmul.allow_tf32 = bool(tf32)
    st = AttnGANStep(ns, B, device=dev, damsm=damsm, init=init or ("fast" if dev.type == "cpu" else "weights_init"))
    for _ in range(warmup):
        st.step()
    if dev.type == "cuda":
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            st.step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
    else:
        t0 = time.perf_counter()
        for _ in range(steps):
            st.step()
        ms = 1e3 * (time.perf_counter() - t0) / steps
    return {"ms_per_step": ms, "images_per_s": B / (ms / 1e3), "batch": B, "steps": steps, "warmup": warmup,
            "device": str(dev), "threads": threads if dev.type == "cpu" else None, "damsm": bool(damsm)}


if __name__ == "__main__":
    import argparse
    import json
    ap = argparse.ArgumentParser()
    ap.add_argument("--device", default="cpu")
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--no-damsm", action="store_true")
    ap.add_argument("--tf32", type=int, default=None)
    a = ap.parse_args()
    print(json.dumps(time_attngan(a.batch, a.steps, a.warmup, a.device, not a.no_damsm, a.tf32)))
