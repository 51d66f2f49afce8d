import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "multiple-objects-gan_b200"), os.path.join(ROOT, "tests")]
import numpy as np, torch
import golden_util as gu
from mog_b200 import synth, ops
from mog_b200.attngan import model as M
from mog_b200.attngan.miscc.config import cfg, reset_cfg
from mog_b200.attngan.trainer import condGANTrainer
from oracle import attngan_oracle as O
prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
torch.backends.cudnn.allow_tf32 = False
G, meta = gu.load("attngan_tiny_trainstep")
c, seed, K = meta["cfg"], meta["seed"], meta["steps"]
reset_cfg()
cfg.GAN.GF_DIM, cfg.GAN.DF_DIM, cfg.GAN.Z_DIM = c["GF_DIM"], c["DF_DIM"], c["Z_DIM"]
cfg.GAN.R_NUM, cfg.TEXT.EMBEDDING_DIM, cfg.TEXT.WORDS_NUM = c["R_NUM"], c["EMBEDDING_DIM"], c["T"]
cfg.TRAIN.BATCH_SIZE = c["B"]
cfg.TRAIN.SMOOTH.GAMMA1, cfg.TRAIN.SMOOTH.GAMMA2, cfg.TRAIN.SMOOTH.GAMMA3, cfg.TRAIN.SMOOTH.LAMBDA = 4.0, 5.0, 10.0, 50.0
cfg.MOG.PRECISION = prec
netG = M.G_NET(); netsD = [M.D_NET64(), M.D_NET128(), M.D_NET256()]
netG.load_state_dict(synth.fill_state_dict(netG.state_dict(), seed + 1))
for i, d in enumerate(netsD):
    d.load_state_dict(synth.soften_logits(synth.fill_state_dict(d.state_dict(), seed + 2 + i), meta["logit_scale"]))
PG = O.leafify(netG.state_dict()); PDs = [O.leafify(d.state_dict()) for d in netsD]
ocfg = O.Cfg(GF_DIM=c["GF_DIM"], DF_DIM=c["DF_DIM"], Z_DIM=c["Z_DIM"], R_NUM=c["R_NUM"], EMBEDDING_DIM=c["EMBEDDING_DIM"])
ostate = O.make_train_state(PG, PDs)
netG.cuda().train()
for d in netsD: d.cuda().train()
tr = condGANTrainer("", None, 0, None)
ops.set_precision(prec)
tr.image_encoder = synth.StandInEncoder(c["EMBEDDING_DIM"], device="cuda")
optG, optDs = tr.define_optimizers(netG, netsD)
st = tr.make_step_state(netG, netsD, optG, optDs)
batch = synth.attngan_batch(c["B"], T=c["T"], nef=c["EMBEDDING_DIM"], nz=c["Z_DIM"], seed=seed)
b = {k: ([t.cuda() for t in v] if isinstance(v, list) else (v.cuda() if torch.is_tensor(v) else v)) for k, v in batch.items()}
oenc = synth.StandInEncoder(c["EMBEDDING_DIM"])
def rel(a, r):
    a = a.detach().cpu().double(); r = r.detach().double()
    return float((a - r).norm() / r.norm().clamp_min(1e-30))
cap = {}
netG.register_forward_hook(lambda m, i, o: cap.__setitem__("imgs", [t.detach() for t in o[0]]))
for k in range(K):
    noise = torch.from_numpy(np.random.RandomState(seed + 10 + k).standard_normal((c["B"], c["Z_DIM"])).astype(np.float32))
    eps = gu.full(G, "step%d/eps" % k)
    out = O.train_step(PG, PDs, ostate, ocfg, batch, eps=eps, noise=noise, image_encoder=oenc)
    errD, errG, kl = tr.train_step(st, b["imgs"], b["sent_emb"], b["words_embs"], b["mask"], b["transf_matrices"], b["transf_matrices_inv"],
                                   b["label_one_hot"], b["cap_lens"], b["class_ids"], noise=noise.cuda(), eps=eps.cuda())
    print("   fake imgs rel:", [rel(cap["imgs"][i], out["fake_imgs"][i]) for i in range(3)])
    print("step", k, "errD", float(errD), float(sum(out["errD"])), "errG", float(errG), float(out["errG"] + out["kl"]))
    rows = []
    for name, p in netG.named_parameters():
        rows.append((rel(p.grad, PG[name].grad), rel(p, PG[name]), "G " + name))
    for i, d in enumerate(netsD):
        for name, p in d.named_parameters():
            rows.append((rel(p.grad, PDs[i][name].grad), rel(p, PDs[i][name]), "D%d %s" % (i, name)))
    for name, v in netG.state_dict().items():
        if "running" in name: rows.append((0.0, rel(v, PG[name]), "Gbuf " + name))
    for g, p, n in rows:
        if n.startswith("G img_net") or n.startswith("D1 img") or n.startswith("D0 conv") or n.startswith("D2 img_code_s16"):
            nm = n.split(" ")[1]
            net = netG if n.startswith("G ") else netsD[int(n[1])]
            gn = float(dict(net.named_parameters())[nm].grad.norm())
            print("   grad %.2e  param %.2e  %-34s |g| %.3e" % (g, p, n, gn))
    rows.sort(key=lambda r: -max(r[0], r[1]))
    for g, p, n in [r for r in rows if r[2].startswith("G ")][:4]:
        print("   TOP-G grad %.2e  param %.2e  %s" % (g, p, n))
    pm = dict(netG.named_parameters())["img_net2.img.0.weight"]
    print("   img_net2 grad norm mine %.4e ref %.4e" % (float(pm.grad.norm()), float(PG["img_net2.img.0.weight"].grad.norm())))
