"""The boundary exercised the way the reference's ``main.py`` uses it (main.py:136-160): ``from trainer import
condGANTrainer`` resolved through ``dropin/attngan`` on sys.path, a ``DataLoader`` over the reference's tuple format, frozen
DAMSM encoders loaded from ``cfg.TRAIN.NET_E``, ``algo.train()`` with no arguments, the checkpoint it writes, then the
sampling entry points on that checkpoint.  Run on the B200 box: -m gpu."""
import glob
import os
import sys

import numpy as np
import pytest
import torch
import torch.utils.data

from mog_b200 import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _tiny_cfg(tmp):
    from mog_b200.attngan.miscc.config import cfg, reset_cfg
    reset_cfg()
    cfg.GAN.GF_DIM, cfg.GAN.DF_DIM, cfg.GAN.Z_DIM, cfg.GAN.R_NUM = 8, 8, 16, 1
    cfg.TEXT.EMBEDDING_DIM, cfg.TEXT.WORDS_NUM = 32, 12
    cfg.TRAIN.BATCH_SIZE, cfg.TRAIN.MAX_EPOCH, cfg.TRAIN.SNAPSHOT_INTERVAL = 2, 1, 1
    cfg.TRAIN.SMOOTH.GAMMA1, cfg.TRAIN.SMOOTH.GAMMA2, cfg.TRAIN.SMOOTH.GAMMA3, cfg.TRAIN.SMOOTH.LAMBDA = 4.0, 5.0, 10.0, 50.0
    cfg.MOG.PRECISION = 'bf16x3'
    return cfg


def _write_encoders(tmp, n_words, nef):
    from mog_b200.attngan import model as M
    te = M.RNN_ENCODER(n_words, nhidden=nef)
    ie = M.CNN_ENCODER(nef)
    ie.load_state_dict(synth.fill_encoder_state_dict(ie.state_dict(), 9))
    path = os.path.join(tmp, "text_encoder100.pth")
    torch.save(te.state_dict(), path)
    torch.save(ie.state_dict(), os.path.join(tmp, "image_encoder100.pth"))
    return path


def test_main_py_call_order_trains_and_samples(tmp_path):
    tmp = str(tmp_path)
    cfg = _tiny_cfg(tmp)
    n_words = 40
    cfg.TRAIN.NET_E = _write_encoders(tmp, n_words, cfg.TEXT.EMBEDDING_DIM)
    sys.path.insert(0, os.path.join(ROOT, "dropin", "attngan"))
    try:
        for m in ("trainer", "model"):
            sys.modules.pop(m, None)
        from trainer import condGANTrainer as trainer      # main.py:5
    finally:
        sys.path.pop(0)
    ds = synth.SyntheticTextDataset(n=4, n_words=n_words, words_num=cfg.TEXT.WORDS_NUM, seed=1)
    dataloader = torch.utils.data.DataLoader(ds, batch_size=cfg.TRAIN.BATCH_SIZE, drop_last=True, shuffle=True, num_workers=0)
    output_dir = os.path.join(tmp, "out")
    algo = trainer(output_dir, dataloader, n_words, ds.ixtoword, False)    # main.py:139
    st = algo.train()                                                      # main.py:152: no arguments
    assert algo.image_encoder is not None and algo.text_encoder is not None   # loaded from cfg.TRAIN.NET_E, DAMSM on
    assert any(k == "w_loss" for k, _ in st["last_logs"])
    for _, v in st["last_logs"]:
        assert bool(torch.isfinite(v))
    ckpts = sorted(glob.glob(os.path.join(output_dir, "Model", "*.pth")))
    assert ckpts, "train() must leave a checkpoint"
    sd = torch.load(ckpts[-1], map_location="cpu")
    assert set(sd) == {"epoch", "netG", "optimG", "netD", "optimD"} and len(sd["netD"]) == 3      # trainer.py:184-190
    # the saved generator is the EMA copy (trainer.py:182-193)
    ema = dict(zip([n for n, _ in st["netG"].named_parameters()], st["avg_param_G"]))
    for n, a in ema.items():
        assert torch.equal(sd["netG"][n], a.cpu()), n

    # ---- sampling entry points on that checkpoint (generator in eval mode)
    cfg.TRAIN.NET_G = ckpts[-1]
    cfg.TRAIN.FLAG = False
    files = algo.sampling("test", num_samples=1)
    assert len(files) == cfg.TRAIN.BATCH_SIZE and all(os.path.isfile(f) for f in files)
    from PIL import Image
    im = np.asarray(Image.open(files[0]))
    assert im.shape == (256, 256, 3)
    dse = synth.SyntheticTextDataset(n=2, n_words=n_words, words_num=cfg.TEXT.WORDS_NUM, seed=2, eval=True)
    algo.data_loader = torch.utils.data.DataLoader(dse, batch_size=cfg.TRAIN.BATCH_SIZE, drop_last=True, shuffle=False)
    rows = algo.sample("test", num_samples=1, draw_bbox=True)
    assert len(rows) == 1 and os.path.isfile(rows[0])
    caps = np.zeros((3, 7), np.int64)
    lens = np.array([7, 5, 2])
    for i, l in enumerate(lens):
        caps[i, :l] = np.arange(1, l + 1)
    out = algo.gen_example({"example": [caps, lens, np.array([2, 0, 1])]})
    assert len(out) == 3 * 3 and all(os.path.isfile(f) for f in out)


def test_eval_mode_generator_uses_running_statistics():
    """``netG.eval()`` (sampling): BatchNorm folds the running statistics; output is deterministic and batch-independent."""
    from mog_b200.attngan import model as M
    _tiny_cfg("")
    netG = M.G_NET()
    netG.load_state_dict(synth.fill_state_dict(netG.state_dict(), 3))
    netG.cuda().eval()
    b = synth.attngan_batch(3, T=9, nef=32, nz=16, seed=5)
    d = {k: v.cuda() for k, v in b.items() if torch.is_tensor(v)}
    d["mask"] = torch.zeros_like(d["mask"])      # (the mask-tiling quirk makes masked positions depend on the batch size)
    with torch.no_grad():
        full, _, _, _ = netG(d["noise"], d["sent_emb"], d["words_embs"], d["mask"], d["transf_matrices_inv"], d["label_one_hot"], eps=d["eps"])
        one, _, _, _ = netG(d["noise"][:1], d["sent_emb"][:1], d["words_embs"][:1], d["mask"][:1], d["transf_matrices_inv"][:1],
                            d["label_one_hot"][:1], eps=d["eps"][:1])
    # sample 0 alone == sample 0 inside a batch (train-mode batch statistics would differ)
    for i in range(3):
        rel = float((full[i][:1] - one[i]).norm() / one[i].norm())
        assert rel < 1e-5, (i, rel)


def test_packed_weights_follow_data_writes():
    """``load_params`` / ``weights_init`` write through ``p.data`` (no version bump); the packed conv operands
    must still be refreshed: forward, load_params, forward == a fresh module with those parameters."""
    from mog_b200 import ops
    from mog_b200.attngan import model as M
    from mog_b200.attngan.miscc.utils import copy_G_params, load_params
    _tiny_cfg("")
    ops.set_precision("bf16x3")
    torch.manual_seed(0)
    d1, d2 = M.D_NET128().cuda().train(), M.D_NET128().cuda().train()
    d1.load_state_dict(synth.fill_state_dict(d1.state_dict(), 1))
    d2.load_state_dict(synth.fill_state_dict(d2.state_dict(), 2))
    x = torch.rand(2, 3, 128, 128, device="cuda") * 2 - 1
    with torch.no_grad():
        y1 = d1(x).clone()
        load_params(d1, copy_G_params(d2))           # d1 now carries d2's parameters
        y12 = d1(x)
        y2 = d2(x)
    assert float((y1 - y2).abs().max()) > 1e-3       # the two nets differ
    assert float((y12 - y2).norm() / y2.norm()) < 1e-6


def test_integration_md_stub_runs():
    """The ctypes stub printed in INTEGRATION.md section 2 is executed as written (argument count / byte sizes of the C ABI)."""
    import re
    import torch.nn.functional as F
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    block = re.search(r"```python\n(import ctypes, torch\n.*?)```", text, re.S).group(1)
    block = block.replace('ctypes.CDLL("mog_b200/libmog.so")',
                          'ctypes.CDLL(%r)' % os.path.join(ROOT, "multiple-objects-gan_b200", "mog_b200", "libmog.so"))
    ns = {}
    exec(compile(block, "INTEGRATION.md", "exec"), ns)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 96, 32, 32, generator=g)
    w = torch.randn(96, 96, 3, 3, generator=g) / 29.4
    y = ns["conv3x3_up2x"](x.permute(0, 2, 3, 1).contiguous().cuda(), w.cuda())
    torch.cuda.synchronize()
    ref = F.conv2d(F.interpolate(x, scale_factor=2, mode="nearest"), w, None, 1, 1)
    rel = float((y.permute(0, 3, 1, 2).cpu().double() - ref.double()).norm() / ref.double().norm())
    assert rel < 5e-5, rel
