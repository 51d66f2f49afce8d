"""Run K train steps (tiny config, fp32 mode) with EVERY Conv2dFn.backward checked against torch fp64 on the SAME inputs."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "multiple-objects-gan_b200"), os.path.join(ROOT, "tests")]
import numpy as np, torch, torch.nn.functional as F
import golden_util as gu
from mog_b200 import synth, ops
from mog_b200.attngan import model as M
from mog_b200.attngan.miscc.config import cfg, reset_cfg
from mog_b200.attngan.trainer import condGANTrainer
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
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
netG.cuda().train()
for d in netsD: d.cuda().train()
names = {}
for tag, net in [("G", netG)] + [("D%d" % i, d) for i, d in enumerate(netsD)]:
    for n, p in net.named_parameters():
        names[id(p)] = tag + "." + n
tr = condGANTrainer("", None, 0, None)
tr.image_encoder = synth.StandInEncoder(c["EMBEDDING_DIM"], device="cuda")
optG, optDs = tr.define_optimizers(netG, netsD)
st = tr.make_step_state(netG, netsD, optG, optDs)
batch = synth.attngan_batch(c["B"], T=c["T"], nef=c["EMBEDDING_DIM"], nz=c["Z_DIM"], seed=seed)
b = {k: ([t.cuda() for t in v] if isinstance(v, list) else (v.cuda() if torch.is_tensor(v) else v)) for k, v in batch.items()}
step_no = [0]
orig_fwd, orig_bwd = ops.Conv2dFn.forward, ops.Conv2dFn.backward
ACT = {0: lambda z: z, 1: F.relu, 2: lambda z: F.leaky_relu(z, 0.2), 4: torch.tanh, 5: torch.sigmoid}
def fwd(ctx, x, weight, bias, stride, pad, up2x, act, precision):
    ctx.dbg = (x.detach(), weight, bias)
    return orig_fwd(ctx, x, weight, bias, stride, pad, up2x, act, precision)
def bwd(ctx, dy):
    out = orig_bwd(ctx, dy)
    x, weight, bias = ctx.dbg
    stride, pad, up2x, act, precision, xshape = ctx.cfg
    if x.dim() != 4: return out
    with torch.enable_grad():
        xd = x.double().permute(0, 3, 1, 2).detach().requires_grad_(True)
        wd = weight.detach().double()
        if wd.dim() == 2: wd = wd.reshape(wd.shape[0], wd.shape[1], 1, 1)
        wd = wd.detach().requires_grad_(True)
        xin = F.interpolate(xd, scale_factor=2, mode="nearest") if up2x else xd
        z = F.conv2d(xin, wd, None if bias is None else bias.detach().double(), stride, pad)
        y = ACT[act](z)
        y.backward(dy.double().permute(0, 3, 1, 2))
    rel = lambda a, r: float((a.double() - r).norm() / r.norm().clamp_min(1e-300))
    msgs = []
    if out[0] is not None:
        e = rel(out[0].permute(0, 3, 1, 2), xd.grad)
        if e > 2e-5: msgs.append("dx %.2e" % e)
    if out[1] is not None:
        e = rel(out[1].reshape(wd.shape), wd.grad)
        if e > 2e-5: msgs.append("dw %.2e" % e)
    if msgs:
        print("step %d conv %-40s x%s w%s s%d p%s up%d act%d: %s" % (step_no[0], names.get(id(weight), "?"), tuple(x.shape), tuple(weight.shape), stride, pad, up2x, act, " ".join(msgs)))
    return out
ops.Conv2dFn.forward = staticmethod(fwd); ops.Conv2dFn.backward = staticmethod(bwd)
for k in range(K):
    step_no[0] = k
    noise = torch.from_numpy(np.random.RandomState(seed + 10 + k).standard_normal((c["B"], c["Z_DIM"])).astype(np.float32))
    eps = gu.full(G, "step%d/eps" % k)
    tr.train_step(st, b["imgs"], b["sent_emb"], b["words_embs"], b["mask"], b["transf_matrices"], b["transf_matrices_inv"],
                  b["label_one_hot"], b["cap_lens"], b["class_ids"], noise=noise.cuda(), eps=eps.cuda())
torch.cuda.synchronize()
print("selfcheck done")
