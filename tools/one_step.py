"""Runs W warm-up + 1 training step of the bench workload (for ncu launch lists)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "multiple-objects-gan_b200")]
import torch
import bench
from mog_b200 import ops, synth
from mog_b200.attngan.trainer import condGANTrainer
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
warm = int(sys.argv[3]) if len(sys.argv) > 3 else 1
cfg = bench.set_cfg(); ops.set_precision(prec); cfg.TRAIN.BATCH_SIZE = B
torch.manual_seed(1234)
tr = condGANTrainer("", None, 0, None)
enc = None
if os.environ.get("MOG_NO_DAMSM", "0") != "1":
    from mog_b200.attngan.model import CNN_ENCODER
    enc = CNN_ENCODER(256)
    enc.load_state_dict(synth.fill_encoder_state_dict(enc.state_dict(), 9))
    for p in enc.parameters():
        p.requires_grad = False
    enc.cuda().eval()
_, _, netG, netsD, _ = tr.build_models(image_encoder=enc, load_encoders=False)
optG, optDs = tr.define_optimizers(netG, netsD)
st = tr.make_step_state(netG, netsD, optG, optDs)
h = synth.attngan_batch(B, seed=1234)
d = {k: v.cuda() for k, v in h.items() if torch.is_tensor(v)}
imgs = [t.cuda() for t in h["imgs"]]
for i in range(warm + 1):
    if i == warm:
        torch.cuda.synchronize(); torch.cuda.profiler.start()
    tr.train_step(st, imgs, d["sent_emb"], d["words_embs"], d["mask"], d["transf_matrices"], d["transf_matrices_inv"], d["label_one_hot"], h["cap_lens"], h["class_ids"])
torch.cuda.synchronize(); torch.cuda.profiler.stop()
print("done")
