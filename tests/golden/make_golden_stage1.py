#!/usr/bin/env python
"""Golden fixtures for the single-stage programs, by EXECUTING THE UNMODIFIED REFERENCE:

    python tests/golden/make_golden_stage1.py mnist
    python tests/golden/make_golden_stage1.py clevr

(one program per process: the four reference programs use the same module names).  Imports
/root/reference/code/{multi-mnist,clevr}/{model.py,miscc/utils.py} through the harness shims of
SURVEY.md section 8(c) and runs G forward, the D loss + backward and the G loss + backward of the
reference training step (multi-mnist/trainer.py:134-157, clevr/trainer.py:130-154) on the
deterministic synthetic batch / weights of mog_b200.synth."""
import json
import os
import sys

import numpy as np
import torch
import torch.nn as nn

prog = sys.argv[1]
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/code/" + {"mnist": "multi-mnist", "clevr": "clevr"}[prog]
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

C = dict(GF_DIM=8, DF_DIM=8, Z_DIM=20, CONDITION_DIM=16, B=4)
cfg.CUDA = False
cfg.GAN.GF_DIM, cfg.GAN.DF_DIM, cfg.Z_DIM, cfg.GAN.CONDITION_DIM = C["GF_DIM"], C["DF_DIM"], C["Z_DIM"], C["CONDITION_DIM"]
cfg.USE_BBOX_LAYOUT = True
seed = 300 if prog == "mnist" else 400
torch.manual_seed(seed)
b = synth.stage1_batch(prog, C["B"], nz=C["Z_DIM"], seed=seed)
netG, netD = M.STAGE1_G(), M.STAGE1_D()
netG.load_state_dict(synth.fill_state_dict(netG.state_dict(), seed + 1))
netD.load_state_dict(synth.fill_state_dict(netD.state_dict(), seed + 2))
netG.train()
netD.train()
E = {}
out = netG(b["noise"], b["transf_matrices_inv"], b["label_one_hot"])
fake = out[1] if isinstance(out, tuple) else out
E["fake"] = summarize(fake)
real_labels, fake_labels = torch.ones(C["B"]), torch.zeros(C["B"])
netD.zero_grad()
errD, _, _, _ = U.compute_discriminator_loss(netD, b["imgs"], fake, real_labels, fake_labels, b["label_one_hot"].clone(),
                                             b["transf_matrices"], b["transf_matrices_inv"], [0])
errD.backward(retain_graph=True)
E["errD"] = summarize(errD)
for k, p in netD.named_parameters():
    E["D/grad/" + k] = summarize(p.grad)
for k, v in netD.state_dict().items():
    if "running" in k:
        E["D/buf/" + k] = summarize(v)
netG.zero_grad()
errG = U.compute_generator_loss(netD, fake, real_labels, b["label_one_hot"].clone(), b["transf_matrices"],
                                b["transf_matrices_inv"], [0])
errG.backward()
E["errG"] = summarize(errG)
for k, p in netG.named_parameters():
    if p.grad is not None:
        E["G/grad/" + k] = summarize(p.grad)
save("stage1_" + prog, E, {"cfg": C, "seed": seed, "program": prog,
                           "what": "reference STAGE1_G fwd, D loss+bwd, G loss+bwd (no optimiser step)"})
keys = {"STAGE1_G": {k: list(v.shape) for k, v in netG.state_dict().items()},
        "STAGE1_D": {k: list(v.shape) for k, v in netD.state_dict().items()}}
with open(os.path.join(HERE, "stage1_%s_keys.json" % prog), "w") as f:
    json.dump(keys, f, indent=0)
print("done", prog)
