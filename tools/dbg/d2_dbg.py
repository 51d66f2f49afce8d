import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "multiple-objects-gan_b200"), os.path.join(ROOT, "tests")]
import numpy as np, torch
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
from baseline import ref_harness as H
ns = H.load("attngan", "cuda")
from mog_b200 import synth, ops
from mog_b200.attngan import model as M
from mog_b200.attngan.miscc import losses as L
from mog_b200.attngan.miscc.config import cfg, reset_cfg
c = dict(GF_DIM=8, DF_DIM=8, Z_DIM=20, R_NUM=2, EMBEDDING_DIM=32, T=6, B=4)
reset_cfg()
cfg.GAN.GF_DIM, cfg.GAN.DF_DIM, cfg.GAN.Z_DIM = c["GF_DIM"], c["DF_DIM"], c["Z_DIM"]
cfg.GAN.R_NUM, cfg.TEXT.EMBEDDING_DIM, cfg.TEXT.WORDS_NUM = c["R_NUM"], c["EMBEDDING_DIM"], c["T"]
H.set_attngan_cfg(ns, c, c["B"])
which = int(sys.argv[1]) if len(sys.argv) > 1 else 2
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 0.02
cls = [M.D_NET64, M.D_NET128, M.D_NET256][which]; rcls = [ns.model.D_NET64, ns.model.D_NET128, ns.model.D_NET256][which]
sd = synth.soften_logits(synth.fill_state_dict(cls().state_dict(), 202 + which), scale)
rng = np.random.RandomState(0)
S = 64 << which
B = c["B"]
real = torch.from_numpy(rng.uniform(-1, 1, (B, 3, S, S)).astype(np.float32)).cuda()
fake = torch.tanh(torch.from_numpy(rng.standard_normal((B, 3, S, S)).astype(np.float32))).cuda()
sent = torch.from_numpy(np.tanh(rng.standard_normal((B, 32))).astype(np.float32)).cuda()
b = synth.attngan_batch(B, T=6, nef=32, nz=20, seed=200)
oh, tm, tmi = b["label_one_hot"].cuda(), b["transf_matrices"].cuda(), b["transf_matrices_inv"].cuda()
ones, zeros = torch.ones(B, device="cuda"), torch.zeros(B, device="cuda")
# reference in fp64
ref = rcls().double().cuda(); ref.load_state_dict({k: v.double() if v.is_floating_point() else v for k, v in sd.items()}); ref.train()
kw64 = dict(local_labels=oh.double(), transf_matrices=tm.double(), transf_matrices_inv=tmi.double()) if which == 0 else {}
e = ns.losses.discriminator_loss(ref, real.double(), fake.double(), sent.double(), ones.double(), zeros.double(), [0], **kw64)
e.backward()
rg = {k: p.grad.clone() for k, p in ref.named_parameters()}
# reference in fp32 (cuDNN, tf32 off)
ref32 = rcls().cuda(); ref32.load_state_dict(sd); ref32.train()
kw32 = dict(local_labels=oh, transf_matrices=tm, transf_matrices_inv=tmi) if which == 0 else {}
e32 = ns.losses.discriminator_loss(ref32, real, fake, sent, ones, zeros, [0], **kw32)
e32.backward()
def run(prec, pair):
    ops.set_precision(prec); L.PAIR_PASS = pair
    net = cls(); net.load_state_dict(sd); net.cuda().train()
    err = L.discriminator_loss(net, real, fake, sent, ones, zeros, [0], **kw32)
    err.backward()
    return float(err), {k: p.grad.clone() for k, p in net.named_parameters()}
outs = {"fp32 pair": run("fp32", True), "fp32 calls": run("fp32", False), "x3 pair": run("bf16x3", True)}
print("loss ref64 %.9f ref32 %.9f" % (float(e), float(e32)), {k: v[0] for k, v in outs.items()})
rel = lambda a, r: float((a.double() - r).norm() / r.norm())
print("%-34s %10s %10s %10s %10s" % ("param", "ref32", *outs.keys()))
for k in rg:
    print("%-34s %10.2e %10.2e %10.2e %10.2e" % (k, rel(dict(ref32.named_parameters())[k].grad, rg[k]), *[rel(v[1][k], rg[k]) for v in outs.values()]))
