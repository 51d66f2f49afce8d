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
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
sd = synth.soften_logits(synth.fill_state_dict(M.D_NET256().state_dict(), 204), 0.02)
rng = np.random.RandomState(0)
B = c["B"]
real = torch.from_numpy(rng.uniform(-1, 1, (B, 3, 256, 256)).astype(np.float32)).cuda()
fake = torch.tanh(torch.from_numpy(rng.standard_normal((B, 3, 256, 256)).astype(np.float32))).cuda()
sent = torch.from_numpy(np.tanh(rng.standard_normal((B, 32))).astype(np.float32)).cuda()
if os.environ.get("GFAKE", "0") == "1":
    import json
    import golden_util as gu
    from oracle import attngan_oracle as O
    keys = json.load(open(os.path.join(ROOT, "tests", "golden", "attngan_state_dict_keys.json")))
    Gd, meta = gu.load("attngan_tiny_trainstep")
    seed = meta["seed"]
    PG = O.leafify(synth.fill_state_dict({k: torch.empty(s_) for k, s_ in keys["tiny"]["G_NET"].items()}, seed + 1))
    bt = synth.attngan_batch(B, T=6, nef=32, nz=20, seed=seed)
    ocfg = O.Cfg(GF_DIM=8, DF_DIM=8, Z_DIM=20, R_NUM=2, EMBEDDING_DIM=32)
    noise = torch.from_numpy(np.random.RandomState(seed + 10).standard_normal((B, 20)).astype(np.float32))
    fk = O.g_net(PG, ocfg, noise, bt["sent_emb"], bt["words_embs"], bt["mask"], bt["transf_matrices_inv"], bt["label_one_hot"], eps=gu.full(Gd, "step0/eps"))[0][2]
    fake = fk.detach().cuda(); real = bt["imgs"][2].cuda(); sent = bt["sent_emb"].cuda()
    sd = synth.soften_logits(synth.fill_state_dict(M.D_NET256().state_dict(), seed + 4), 0.02)
    print("fake: std over pixels per sample", fake.std(dim=(1, 2, 3)).tolist(), " std across samples", float(fake.std(dim=0).mean()), "mean", float(fake.mean()))
ones, zeros = torch.ones(B, device="cuda"), torch.zeros(B, device="cuda")
names = ["img_code_s16", "img_code_s32", "img_code_s64", "img_code_s64_1", "img_code_s64_2"]
def tap(net, store):
    for n in names:
        def hook(mod, inp, out, n=n):
            out.retain_grad(); store.setdefault(n, []).append(out)
        getattr(net, n).register_forward_hook(hook)
ref = ns.model.D_NET256().double().cuda(); ref.load_state_dict({k: v.double() if v.is_floating_point() else v for k, v in sd.items()}); ref.train()
R = {}; tap(ref, R)
e = ns.losses.discriminator_loss(ref, real.double(), fake.double(), sent.double(), ones.double(), zeros.double(), [0]); e.backward()
ops.set_precision(prec)
if os.environ.get("DIRTY", "0") == "1":
    # replicate the allocator state of a train step: run one full step on throw-away nets first
    from mog_b200.attngan.trainer import condGANTrainer
    cfg.MOG.PRECISION = prec
    cfg.TRAIN.BATCH_SIZE = B
    cfg.TRAIN.SMOOTH.GAMMA1, cfg.TRAIN.SMOOTH.GAMMA2, cfg.TRAIN.SMOOTH.GAMMA3, cfg.TRAIN.SMOOTH.LAMBDA = 4.0, 5.0, 10.0, 50.0
    tr = condGANTrainer("", None, 0, None)
    tr.image_encoder = synth.StandInEncoder(32, device="cuda")
    nG = M.G_NET(); nDs = [M.D_NET64(), M.D_NET128(), M.D_NET256()]
    nG.load_state_dict(synth.fill_state_dict(nG.state_dict(), seed + 1))
    for i_, d_ in enumerate(nDs): d_.load_state_dict(synth.soften_logits(synth.fill_state_dict(d_.state_dict(), seed + 2 + i_), 0.02))
    nG.cuda().train(); [d_.cuda().train() for d_ in nDs]
    oG, oDs = tr.define_optimizers(nG, nDs); st_ = tr.make_step_state(nG, nDs, oG, oDs)
    bb = {k: ([t.cuda() for t in v] if isinstance(v, list) else (v.cuda() if torch.is_tensor(v) else v)) for k, v in bt.items()}
    snap = {}
    orig_opt = tr._opt_step
    def spy(opt, grad_scale=1.0):
        if opt is oDs[2]:
            torch.cuda.synchronize()
            snap.update({k: p.grad.clone() for k, p in nDs[2].named_parameters()})
        return orig_opt(opt, grad_scale)
    tr._opt_step = spy
    tr.train_step(st_, bb["imgs"], bb["sent_emb"], bb["words_embs"], bb["mask"], bb["transf_matrices"], bb["transf_matrices_inv"], bb["label_one_hot"], bb["cap_lens"], bb["class_ids"], noise=noise.cuda(), eps=gu.full(Gd, "step0/eps").cuda())
    g2 = {k: p.grad.clone() for k, p in nDs[2].named_parameters()}
    ops.set_precision(prec)
L.PAIR_PASS = os.environ.get("PAIR", "1") == "1"
net = M.D_NET256(); net.load_state_dict(sd); net.cuda().train()
T = {}; tap(net, T)
err = L.discriminator_loss(net, real, fake, sent, ones, zeros, [0]); err.backward()
rel = lambda a, r: float((a.double() - r).norm() / r.norm())
for n in names:
    if len(T[n]) == 1:
        mine = T[n][0]                       # [2B, H, W, C] NHWC (block output)
        m_act = mine.detach().permute(0, 3, 1, 2); m_g = mine.grad.permute(0, 3, 1, 2)
    else:
        m_act = torch.cat([t.detach() for t in T[n][:2]], 0).permute(0, 3, 1, 2); m_g = torch.cat([t.grad for t in T[n][:2]], 0).permute(0, 3, 1, 2)
    r_act = torch.cat((R[n][0].detach(), R[n][1].detach()), 0); r_g = torch.cat((R[n][0].grad, R[n][1].grad), 0)
    rf = r_act[B:]
    print("%-16s act %.2e   grad real %.2e fake %.2e   | fake seg: max over ch of |mean|/std %.1f" % (n, rel(m_act, r_act), rel(m_g[:B], r_g[:B]), rel(m_g[B:], r_g[B:]),
          float((rf.mean(dim=(0, 2, 3)).abs() / rf.std(dim=(0, 2, 3))).max())))
    if n == "img_code_s64_1":
        d = (m_g.double() - r_g)
        print("   grad err: per-channel mean of err / rms err:", float(d.mean(dim=(0, 2, 3)).abs().mean() / d.pow(2).mean().sqrt()),
              " max abs err / rms grad", float(d.abs().max() / r_g.pow(2).mean().sqrt()), " n big", int((d.abs() > 0.01 * r_g.pow(2).mean().sqrt()).sum()), "of", d.numel())

if os.environ.get("DIRTY", "0") == "1":
    print("grads: trainstep-D2 vs ref64 | isolated-D2 (after dirtying) vs ref64")
    rg = dict(ref.named_parameters())
    for k, p in net.named_parameters():
        if "s16" in k or "s32.0" in k:
            print("  %-30s %.2e   %.2e   at-Adam-time %.2e  same-object-changed %s" % (k, rel(g2[k], rg[k].grad), rel(p.grad, rg[k].grad), rel(snap[k], rg[k].grad), bool((snap[k] != g2[k]).any())))
