#!/usr/bin/env python
"""Golden fixtures for the DAMSM image encoder, by EXECUTING THE UNMODIFIED REFERENCE ``CNN_ENCODER``
(/root/reference/code/coco/attngan/model.py:207-313) on top of the container's torchvision Inception-v3:

    python tests/golden/make_golden_encoder.py

Shim 7 of SURVEY.md section 8(c): ``model_zoo.load_url`` (no network) returns a deterministic stand-in for the
ImageNet weights (``mog_b200.synth.fill_encoder_state_dict``).  The encoder runs frozen in eval() mode like in
``trainer.py:71-77``; stored: region features, cnn_code, and the gradient of a fixed scalar projection of
both w.r.t. the input image (what ``generator_loss`` back-propagates into the generator)."""
import json
import os
import sys

import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/code/coco/attngan"
sys.path[:0] = [os.path.join(ROOT, "baseline", "shims"), REF, os.path.join(ROOT, "multiple-objects-gan_b200"), os.path.join(ROOT, "tests")]
torch.cuda.FloatTensor = torch.FloatTensor

from mog_b200 import synth  # noqa: E402
from golden_util import save, summarize  # noqa: E402
import torch.utils.model_zoo as model_zoo  # noqa: E402
import torchvision  # noqa: E402

SEED, NEF, B = 700, 32, 2
_full = torchvision.models.inception_v3(weights=None, aux_logits=True, init_weights=False).state_dict()
with open(os.path.join(HERE, "inception_full_shapes.json"), "w") as f:   # the tests rebuild the same weights from these shapes
    json.dump({k: list(v.shape) for k, v in _full.items()}, f, indent=0)
_full = synth.fill_encoder_state_dict(_full, SEED)
model_zoo.load_url = lambda url, *a, **k: _full

from miscc.config import cfg  # noqa: E402  (reference)
import model as M  # noqa: E402  (reference, unmodified)

cfg.CUDA = False
cfg.TRAIN.FLAG = True
enc = M.CNN_ENCODER(NEF)
own = synth.fill_encoder_state_dict({k: v for k, v in enc.state_dict().items() if k.startswith("emb_")}, SEED + 1)
enc.load_state_dict(dict(enc.state_dict(), **own))
for p in enc.parameters():
    p.requires_grad = False
enc.eval()
img, pf, pc = synth.encoder_probe(B, NEF, SEED + 2)
img.requires_grad_(True)
feat, code = enc(img)
loss = (feat * pf).sum() + (code * pc).sum()
loss.backward()
E = {"features": summarize(feat), "cnn_code": summarize(code), "loss": summarize(loss), "d_img": summarize(img.grad)}
save("cnn_encoder", E, {"seed": SEED, "nef": NEF, "B": B,
                        "what": "reference CNN_ENCODER (eval, frozen) fwd + gradient of a fixed projection w.r.t. the image"})
with open(os.path.join(HERE, "cnn_encoder_keys.json"), "w") as f:
    json.dump({k: list(v.shape) for k, v in enc.state_dict().items()}, f, indent=0)
print("done", float(loss))
