import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "multiple-objects-gan_b200"), os.path.join(ROOT, "tests")]
import numpy as np, torch
import golden_util as gu
from mog_b200 import synth, ops
from mog_b200.attngan import model as M
from mog_b200.attngan.miscc.config import cfg, reset_cfg
ops.set_precision("fp32")
G, meta = gu.load("attngan_tiny_step")
c, seed = meta["cfg"], meta["seed"]
reset_cfg()
cfg.GAN.GF_DIM, cfg.GAN.DF_DIM, cfg.GAN.Z_DIM = c["GF_DIM"], c["DF_DIM"], c["Z_DIM"]
cfg.GAN.R_NUM, cfg.TEXT.EMBEDDING_DIM, cfg.TEXT.WORDS_NUM = c["R_NUM"], c["EMBEDDING_DIM"], c["T"]
netG = M.G_NET(); netG.load_state_dict(synth.fill_state_dict(netG.state_dict(), seed + 1)); netG.cuda().train()
b = synth.attngan_batch(c["B"], T=c["T"], nef=c["EMBEDDING_DIM"], nz=c["Z_DIM"], seed=seed)
d = {k: v.cuda() for k, v in b.items() if torch.is_tensor(v)}
eps = gu.full(G, "G/eps").cuda()
imgs, atts, mu, logvar = netG(d["noise"], d["sent_emb"], d["words_embs"], d["mask"], d["transf_matrices_inv"], d["label_one_hot"], eps=eps)
for k, v in netG.state_dict().items():
    if "running" in k:
        s = G["G/buf/" + k]
        a = v.detach().cpu().double().numpy().reshape(-1)
        ref = s["full"] if "full" in s else s["sample"]
        got = a if "full" in s else a[gu._idx(a.size)]
        e = gu.rel_l2(got, ref)
        if e > 2e-5: print("%-50s %.3e" % (k, e), got[:4], ref[:4])
print("done")
