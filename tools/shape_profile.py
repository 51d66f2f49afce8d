"""Per-call device time of one training step, grouped by (libmog entry, problem shape).

    python tools/shape_profile.py [precision] [B] > gpurun_out/shape_profile.txt

Every libmog call of ONE step (after warm-up) is bracketed by CUDA events on the launching stream; calls
are grouped by entry point + conv descriptor, with the algorithmic FLOPs of the convolutions, so the table
shows which layers run far below the tensor roofline.  (Events between every call serialise nothing on a
single stream; the sum over all rows ~ the step's device time minus torch glue.)"""
import collections
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "multiple-objects-gan_b200")]
import torch  # noqa: E402
import bench  # noqa: E402
from mog_b200 import _lib, ops, optim as mog_optim, synth  # noqa: E402
from mog_b200.attngan.trainer import condGANTrainer  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
cfg = bench.set_cfg()
ops.set_precision(prec)
cfg.TRAIN.BATCH_SIZE = B
torch.manual_seed(1234)
tr = condGANTrainer("", None, 0, None)
from mog_b200.attngan.model import CNN_ENCODER  # noqa: E402
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


def step():
    return tr.train_step(st, imgs, d["sent_emb"], d["words_embs"], d["mask"], d["transf_matrices"], d["transf_matrices_inv"],
                         d["label_one_hot"], h["cap_lens"], h["class_ids"])


for _ in range(2):
    step()
torch.cuda.synchronize()

records = []
orig_call = _lib.call
phase = ["?"]


def desc_key(dp):
    d = C.cast(dp, C.POINTER(_lib.MogConvDesc)).contents
    return (d.N, d.H, d.W, d.Cin, d.Cout, d.KH, d.KW, d.stride, d.pad, d.up2x, d.act, d.pad_w1)


def prof_call(name, *args):
    key = name
    extra = None
    if name in ("mog_conv2d_fwd", "mog_conv2d_dgrad", "mog_conv2d_wgrad"):
        extra = desc_key(args[0])
    elif name in ("mog_bn_stats",):
        extra = tuple(args[1:4])
    elif name in ("mog_affine_act_fwd_planes",):
        extra = tuple(args[7:11])
    elif name in ("mog_bn_act_bwd_reduce",):
        extra = tuple(args[6:10])
    elif name in ("mog_bn_act_bwd_apply_planes",):
        extra = tuple(args[8:12])
    elif name in ("mog_split_planes",):
        extra = tuple(args[1:3])
    elif name in ("mog_split_planes_act",):
        extra = tuple(args[3:5])
    elif name in ("mog_pool2d_fwd", "mog_pool2d_bwd"):
        extra = tuple(a for a in args if isinstance(a, int))[-9:-1]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rc = orig_call(name, *args)
    e1.record()
    records.append((phase[0], key, extra, e0, e1))
    return rc


_lib.call = ops.call = mog_optim.call = prof_call
e_all0, e_all1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e_all0.record()
step()
e_all1.record()
torch.cuda.synchronize()
_lib.call = ops.call = mog_optim.call = orig_call

agg = collections.OrderedDict()
for ph, key, extra, e0, e1 in records:
    k = (key, extra)
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += e0.elapsed_time(e1)


def conv_flops(name, e):
    N, H, W, Ci, Co, KH, KW, s, p, up, act, pw1 = e
    Hi, Wi = (H << up), (W << up)
    pw = p if pw1 == 0 else pw1 - 1
    Ho = (Hi + 2 * p - KH) // s + 1
    Wo = (Wi + 2 * pw - KW) // s + 1
    return 2.0 * N * Ho * Wo * Co * Ci * KH * KW


tot = sum(a[1] for a in agg.values())
print("step device time (events around the step): %.2f ms; sum over libmog calls: %.2f ms; %d calls" %
      (e_all0.elapsed_time(e_all1), tot, len(records)))
byname = collections.defaultdict(float)
for (name, e), (n, ms) in agg.items():
    byname[name] += ms
print("\n== by entry point ==")
for name, ms in sorted(byname.items(), key=lambda x: -x[1]):
    print("%8.3f ms  %5.1f%%  %s" % (ms, 100 * ms / tot, name))
print("\n== convolutions by shape (N,H,W,Cin,Cout,KH,KW,stride,pad,up2x,act,pad_w1) ==")
rows = []
for (name, e), (n, ms) in agg.items():
    if name.startswith("mog_conv2d"):
        fl = conv_flops(name, e) * n
        rows.append((ms, n, name[11:], e, fl / (ms * 1e-3) / 1e12))
for ms, n, name, e, tf in sorted(rows, key=lambda x: -x[0]):
    print("%8.3f ms  x%-3d %-6s %-58s %7.1f TFLOP/s alg" % (ms, n, name, str(e), tf))
print("\n== other calls by shape ==")
rows = [(ms, n, name, e) for (name, e), (n, ms) in agg.items() if not name.startswith("mog_conv2d")]
for ms, n, name, e in sorted(rows, key=lambda x: -x[0])[:60]:
    print("%8.3f ms  x%-3d %-30s %s" % (ms, n, name, str(e)))
