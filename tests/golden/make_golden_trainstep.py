#!/usr/bin/env python
"""Golden fixtures for the full training step (a20), by EXECUTING THE UNMODIFIED REFERENCE:

    python tests/golden/make_golden_trainstep.py

Runs K = 3 consecutive iterations of ``code/coco/attngan/trainer.py:294-342`` (G forward; per D: zero_grad,
``discriminator_loss``, backward, ``optim.Adam(betas=(0.5, 0.999)).step()``; ``generator_loss`` incl. the DAMSM
words / sentence branch, ``KL_loss``, backward, Adam, EMA of the G parameters) with the reference's own modules and
losses through ``baseline/ref_harness.AttnGANStep`` at the tiny configuration of ``make_golden.py``, and stores
after the last step: every parameter of G and the three Ds, the EMA copy, Adam's ``exp_avg`` / ``exp_avg_sq``,
the BatchNorm buffers; and per step the three losses.  The DAMSM image encoder is the fixed differentiable
stand-in of ``mog_b200.synth.StandInEncoder`` (the real Inception-v3 composition is pinned by
``make_golden_config5.py``); noise and the CA_NET eps draw of every step are recorded so that the other
implementations consume the same numbers."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
from baseline import ref_harness as H  # noqa: E402
from golden_util import save, summarize  # noqa: E402

TINY = dict(GF_DIM=8, DF_DIM=8, Z_DIM=20, R_NUM=2, EMBEDDING_DIM=32, T=6, B=4)
SEED, K, LOGIT_SCALE = 200, 3, 0.02


def noise_of(k, B, nz):
    return torch.from_numpy(np.random.RandomState(SEED + 10 + k).standard_normal((B, nz)).astype(np.float32))


if __name__ == "__main__":
    import warnings
    warnings.filterwarnings("ignore")
    torch.set_num_threads(8)
    ns = H.load("attngan", "cpu")
    from mog_b200 import synth
    B = TINY["B"]
    st = H.AttnGANStep(ns, B, c=TINY, seed=SEED, device="cpu", damsm=True, encoder=synth.StandInEncoder(TINY["EMBEDDING_DIM"]),
                       init="fill", logit_scale=LOGIT_SCALE)
    E = {}
    for k in range(K):
        # the reference draws eps from the global generator inside CA_NET (model.py:338): seed, record the draw, re-seed
        torch.manual_seed(SEED + k)
        eps = torch.FloatTensor(B, 100).normal_()
        torch.manual_seed(SEED + k)
        errD, errG, kl, _ = st.step(noise=noise_of(k, B, TINY["Z_DIM"]))
        E["step%d/eps" % k] = summarize(eps)
        E["step%d/errD_total" % k], E["step%d/errG_total" % k], E["step%d/kl" % k] = summarize(errD), summarize(errG), summarize(kl)
    nets = {"G": (st.netG, st.optimizerG)}
    for i, d in enumerate(st.netsD):
        nets["D%d" % i] = (d, st.optimizersD[i])
    for tag, (net, opt) in nets.items():
        for name, p in net.named_parameters():
            E["%s/param/%s" % (tag, name)] = summarize(p)
            s = opt.state[p]
            E["%s/exp_avg/%s" % (tag, name)] = summarize(s["exp_avg"])
            E["%s/exp_avg_sq/%s" % (tag, name)] = summarize(s["exp_avg_sq"])
        for name, v in net.state_dict().items():
            if "running" in name or "num_batches" in name:
                E["%s/buf/%s" % (tag, name)] = summarize(v.float())
    for (name, _), a in zip(st.netG.named_parameters(), st.avg_param_G):
        E["G/ema/%s" % name] = summarize(a)
    save("attngan_tiny_trainstep", E, {"cfg": TINY, "seed": SEED, "steps": K, "logit_scale": LOGIT_SCALE,
                                       "what": "reference trainer.py:294-342 x3 (Adam + EMA, DAMSM via StandInEncoder)"})
