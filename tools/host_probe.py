"""Host enqueue time vs device time of segments of the step (is the GPU waiting for Python?)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "multiple-objects-gan_b200")]
import torch
import bench
from mog_b200 import ops, synth, _lib
from mog_b200.attngan.trainer import condGANTrainer
from mog_b200.attngan.model import CNN_ENCODER
from mog_b200.attngan.miscc.losses import discriminator_loss
cfg = bench.set_cfg(); ops.set_precision("bf16x3"); B = 32; cfg.TRAIN.BATCH_SIZE = B
torch.manual_seed(1)
tr = condGANTrainer("", None, 0, None)
enc = CNN_ENCODER(256); enc.load_state_dict(synth.fill_encoder_state_dict(enc.state_dict(), 9))
for p in enc.parameters(): p.requires_grad = False
enc.cuda().eval()
_, _, netG, netsD, _ = tr.build_models(image_encoder=enc, load_encoders=False)
h = synth.attngan_batch(B, seed=1234)
d = {k: v.cuda() for k, v in h.items() if torch.is_tensor(v)}
imgs = [t.cuda() for t in h["imgs"]]
def seg(name, fn, n=3):
    for _ in range(2): out = fn()
    torch.cuda.synchronize()
    hs, ds, ls = [], [], []
    for _ in range(n):
        torch.cuda.synchronize(); l0 = _lib.launch_count(); t0 = time.perf_counter(); out = fn(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
        hs.append(1e3 * (t1 - t0)); ds.append(1e3 * (t2 - t0)); ls.append(_lib.launch_count() - l0)
    print("%-28s host enqueue %7.2f ms   until done %7.2f ms   libmog launches %d" % (name, min(hs), min(ds), ls[0]))
    return out
noise = d["noise"]
def gfwd():
    return netG(noise, d["sent_emb"], d["words_embs"], d["mask"], d["transf_matrices_inv"], d["label_one_hot"])
fake, _, mu, logvar = seg("G forward", gfwd)
ones, zeros = torch.ones(B, device="cuda"), torch.zeros(B, device="cuda")
def dstep(i):
    def f():
        netsD[i].zero_grad(set_to_none=True)
        kw = dict(local_labels=d["label_one_hot"], transf_matrices=d["transf_matrices"], transf_matrices_inv=d["transf_matrices_inv"]) if i == 0 else {}
        e = discriminator_loss(netsD[i], imgs[i], fake[i], d["sent_emb"], ones, zeros, None, **kw); e.backward(); return e
    return f
for i in range(3): seg("D%d loss+backward" % i, dstep(i))
def encfb():
    x = fake[2].detach().requires_grad_(True)
    f, c = enc(x); (f.sum() + c.sum()).backward(); return x.grad
seg("encoder fwd+bwd", encfb)
def gbwd():
    fk, _, mu, lv = gfwd(); (fk[0].sum() + fk[1].sum() + fk[2].sum()).backward()
seg("G forward+backward", gbwd)
