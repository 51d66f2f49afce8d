#!/usr/bin/env python
"""Golden fixtures for the COCO StackGAN program, by EXECUTING THE UNMODIFIED REFERENCE:

    python tests/golden/make_golden_stackgan.py 1      # stage I  (STAGE1_G / STAGE1_D)
    python tests/golden/make_golden_stackgan.py 2      # stage II (STAGE2_G(STAGE1_G) / STAGE2_D)

Imports /root/reference/code/coco/stackgan/{model.py,miscc/utils.py} through the harness shims of
SURVEY.md section 8(c) and runs G forward, the D loss + backward and the G loss (+ KL) + backward of
the reference training step (stackgan/trainer.py:193-235) on the deterministic synthetic batch /
weights of mog_b200.synth.  The CA_NET noise is the reference's own draw
(``torch.FloatTensor(size).normal_()`` under ``torch.manual_seed``); it is recovered by replaying the
generator and stored with the vectors so the oracle and the CUDA path can inject the same eps.

Stage II keeps GF_DIM=192 / CONDITION_DIM=128: the reference hard-codes 768 and 128
(model.py:340,419)."""
import json
import os
import sys

import torch
import torch.nn as nn

stage = int(sys.argv[1])
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/code/coco/stackgan"
sys.path[:0] = [os.path.join(ROOT, "baseline", "shims"), REF, os.path.join(ROOT, "multiple-objects-gan_b200"), os.path.join(ROOT, "tests")]
torch.cuda.FloatTensor = torch.FloatTensor
torch.cuda.DoubleTensor = torch.DoubleTensor


def _dp(module, inputs, device_ids=None, **kw):
    return module(*inputs) if isinstance(inputs, tuple) else module(inputs)


nn.parallel.data_parallel = _dp

from miscc.config import cfg  # noqa: E402  (reference)
import model as M  # noqa: E402  (reference, unmodified)
from miscc import utils as U  # noqa: E402
from mog_b200 import synth  # noqa: E402
from golden_util import save, summarize  # noqa: E402

if stage == 1:
    C = dict(GF_DIM=8, DF_DIM=8, Z_DIM=20, CONDITION_DIM=128, T_DIM=40, R_NUM=1, B=4)
else:
    C = dict(GF_DIM=192, DF_DIM=8, Z_DIM=20, CONDITION_DIM=128, T_DIM=40, R_NUM=1, B=4)
cfg.CUDA = False
cfg.STAGE = stage
cfg.GAN.GF_DIM, cfg.GAN.DF_DIM, cfg.Z_DIM, cfg.GAN.CONDITION_DIM = C["GF_DIM"], C["DF_DIM"], C["Z_DIM"], C["CONDITION_DIM"]
cfg.GAN.R_NUM, cfg.TEXT.DIMENSION = C["R_NUM"], C["T_DIM"]
cfg.USE_BBOX_LAYOUT = True
seed = 500 + stage
B = C["B"]
b = synth.stackgan_batch(B, stage=stage, t_dim=C["T_DIM"], nz=C["Z_DIM"], seed=seed)
if stage == 1:
    netG, netD = M.STAGE1_G(), M.STAGE1_D()
else:
    netG, netD = M.STAGE2_G(M.STAGE1_G()), M.STAGE2_D()
netG.load_state_dict(synth.fill_state_dict(netG.state_dict(), seed + 1))
netD.load_state_dict(synth.fill_state_dict(netD.state_dict(), seed + 2))
netG.train()
netD.train()
E = {}
# the CA_NET draws of the forward below, replayed from the same generator state
torch.manual_seed(seed)
E["eps1"] = summarize(torch.FloatTensor(B, C["CONDITION_DIM"]).normal_())
E["eps2"] = summarize(torch.FloatTensor(B, C["CONDITION_DIM"]).normal_())
torch.manual_seed(seed)
if stage == 1:
    _, fake, mu, logvar, local_labels = netG(b["txt_embedding"], b["noise"], b["transf_matrices_inv"], b["label_one_hot"])
    th, thi = b["transf_matrices"], b["transf_matrices_inv"]
else:
    s1_img, fake, mu, logvar, local_labels = netG(b["txt_embedding"], b["noise"], b["transf_matrices_inv"],
                                                  b["transf_matrices_s2"], b["transf_matrices_inv_s2"], b["label_one_hot"])
    E["stage1_img"] = summarize(s1_img)
    th, thi = b["transf_matrices_s2"], b["transf_matrices_inv_s2"]
E["fake"], E["mu"], E["logvar"], E["local_labels"] = summarize(fake), summarize(mu), summarize(logvar), summarize(local_labels)
real_labels, fake_labels = torch.ones(B), torch.zeros(B)
netD.zero_grad()
errD, _, _, _ = U.compute_discriminator_loss(netD, b["imgs"], fake, real_labels, fake_labels, b["label_one_hot"].clone(), th, thi,
                                             mu, [0])
errD.backward(retain_graph=True)
E["errD"] = summarize(errD)
for k, p in netD.named_parameters():
    E["D/grad/" + k] = summarize(p.grad)
for k, v in netD.state_dict().items():
    if "running" in k:
        E["D/buf/" + k] = summarize(v)
netG.zero_grad()
errG = U.compute_generator_loss(netD, fake, real_labels, b["label_one_hot"].clone(), th, thi, mu, [0])
kl = U.KL_loss(mu, logvar)
(errG + kl * cfg.TRAIN.COEFF.KL).backward()
E["errG"], E["kl"] = summarize(errG), summarize(kl)
for k, p in netG.named_parameters():
    if p.grad is not None:
        E["G/grad/" + k] = summarize(p.grad)
for k, v in netG.state_dict().items():
    if "running" in k:
        E["G/buf/" + k] = summarize(v)
save("stackgan_s%d" % stage, E, {"cfg": C, "seed": seed, "stage": stage, "kl_coeff": cfg.TRAIN.COEFF.KL,
                                 "what": "reference STAGE%d_G fwd, D loss+bwd, G loss+KL bwd (no optimiser step)" % stage})
keys = {"G": {k: list(v.shape) for k, v in netG.state_dict().items()},
        "D": {k: list(v.shape) for k, v in netD.state_dict().items()}}
with open(os.path.join(HERE, "stackgan_s%d_keys.json" % stage), "w") as f:
    json.dump(keys, f, indent=0)
print("done stage", stage)
