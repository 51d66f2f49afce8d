#!/usr/bin/env python
"""Regenerates the algorithmic MAC counts of SURVEY.md section 8(a.1)/8(d) by forward hooks on the UNMODIFIED reference
modules (config 5: GF 48, DF 96, T 18; runs in the build container only -- /root/reference is not on the GPU box):

    python tools/count_macs.py            # writes profiles/macs_config5.json

Counts multiply-accumulates of every nn.Conv2d / nn.Linear call per sample (forward), for G_NET, D_NET64/128/256 (features
+ heads) and CNN_ENCODER, plus the two attention bmm's and the DAMSM word-region products, and derives the per-step totals
bench.py uses (`GMAC_PER_IMAGE_GD`, `GMAC_PER_IMAGE_FULL`): backward of a conv = dgrad + wgrad = 2x forward unless an
operand needs no gradient (first-layer dgrads, frozen D / encoder weights in the G step)."""
import collections
import json
import os
import sys

import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference/code/coco/attngan"
sys.path[:0] = [os.path.join(ROOT, "baseline", "shims"), REF, os.path.join(ROOT, "multiple-objects-gan_b200")]
torch.cuda.FloatTensor = torch.FloatTensor
nn.parallel.data_parallel = lambda m, i, d=None, **k: m(*i) if isinstance(i, tuple) else m(i)
import torch.utils.model_zoo as model_zoo  # noqa: E402
import torchvision  # noqa: E402
model_zoo.load_url = lambda url, *a, **k: torchvision.models.inception_v3(weights=None, aux_logits=True, init_weights=False).state_dict()

from miscc.config import cfg  # noqa: E402  (reference)
import model as M  # noqa: E402  (reference, unmodified)
from mog_b200 import synth  # noqa: E402

cfg.CUDA = False
cfg.TRAIN.FLAG = True
cfg.GAN.GF_DIM, cfg.GAN.DF_DIM, cfg.GAN.Z_DIM, cfg.GAN.R_NUM, cfg.GAN.CONDITION_DIM = 48, 96, 100, 3, 100
cfg.TEXT.EMBEDDING_DIM, cfg.TEXT.WORDS_NUM, cfg.TREE.BRANCH_NUM = 256, 18, 3
B = 2


class Counter:
    def __init__(self):
        self.macs = collections.OrderedDict()
        self.first = {}

    def hook(self, net_name, net):
        def f(mod, inp, out):
            x = inp[0]
            if isinstance(mod, nn.Conv2d):
                m = out.numel() // out.shape[0] * (mod.in_channels // mod.groups) * mod.kernel_size[0] * mod.kernel_size[1]
            else:
                m = out.numel() // out.shape[0] * mod.in_features
            self.macs[net_name] = self.macs.get(net_name, 0) + m * out.shape[0] / B
            # image-fed layers of the discriminators (conv1 / img_code_s16.0 on the 3-channel image, D_NET64.local on the
            # 3 + 81-channel crop): in the D step their input needs no gradient, so their dgrad is not part of the work
            if isinstance(mod, nn.Conv2d) and mod.in_channels in (3, 84):
                self.first[net_name] = self.first.get(net_name, 0) + m * out.shape[0] / B
        hs = [m.register_forward_hook(f) for m in net.modules() if isinstance(m, (nn.Conv2d, nn.Linear))]
        return hs


c = Counter()
b = synth.attngan_batch(B, seed=1)
with torch.no_grad():
    G = M.G_NET().train()
    c.hook("G", G)
    fake, att, mu, logvar = G(b["noise"], b["sent_emb"], b["words_embs"], b["mask"], b["transf_matrices_inv"], b["label_one_hot"])
    Ds = [M.D_NET64().train(), M.D_NET128().train(), M.D_NET256().train()]
    for i, D in enumerate(Ds):
        hs = c.hook("D%d.features" % i, D)
        feat = D(fake[i], b["label_one_hot"], b["transf_matrices"], b["transf_matrices_inv"]) if i == 0 else D(fake[i])
        for h in hs:
            h.remove()
        hs = c.hook("D%d.cond_head" % i, D.COND_DNET)
        D.COND_DNET(feat, b["sent_emb"])
        for h in hs:
            h.remove()
        hs = c.hook("D%d.uncond_head" % i, D.UNCOND_DNET)
        D.UNCOND_DNET(feat)
        for h in hs:
            h.remove()
    E = M.CNN_ENCODER(256).eval()
    c.hook("CNN_ENCODER", E)
    E(fake[2])
g = {k: v / 1e9 for k, v in c.macs.items()}
T, nef, R = 18, 256, 17 * 17
attn = sum(2 * 48 * T * q for q in (64 * 64, 128 * 128)) / 1e9               # two bmm per attention block
Bq = 32                                                                       # DAMSM compares every image with every caption
damsm_fwd = Bq * 2 * R * nef * T / 1e9                                        # per image: B captions x (scores + weighted context)
feat = sum(g["D%d.features" % i] for i in range(3))
cond = g["D0.cond_head"]
first_dgrad = sum(c.first.get("D%d.features" % i, 0) for i in range(3)) / 1e9
out = {
    "forward_gmac_per_sample": dict(g, attention_bmm=attn, damsm_words_B32=damsm_fwd),
    "G_fwd": g["G"] + attn,
    "D_features_sum": feat,
    "cond_head": cond,
}
d_fwd = 2 * feat + 9 * cond
d_bwd = 2 * d_fwd - 2 * first_dgrad      # real + fake pass
g_d_fwd = feat + 3 * cond
out["D_step_fwd"], out["D_step_bwd"] = d_fwd, d_bwd
out["G_step_D_fwd"], out["G_step_D_dgrad"], out["G_bwd"] = g_d_fwd, g_d_fwd, 2 * out["G_fwd"]
out["GMAC_PER_IMAGE_GD"] = out["G_fwd"] + d_fwd + d_bwd + 2 * g_d_fwd + out["G_bwd"]
out["GMAC_PER_IMAGE_FULL"] = out["GMAC_PER_IMAGE_GD"] + 2 * g["CNN_ENCODER"] + 3 * damsm_fwd
os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
with open(os.path.join(ROOT, "profiles", "macs_config5.json"), "w") as f:
    json.dump(out, f, indent=1)
print(json.dumps({k: (round(v, 4) if not isinstance(v, dict) else {a: round(b_, 4) for a, b_ in v.items()}) for k, v in out.items()}, indent=1))
