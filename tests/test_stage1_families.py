"""Multi-MNIST and CLEVR single-stage programs: oracle pinned against the executed reference (CPU),
state_dict contract (CPU), and libmog parity against the golden vectors (GPU)."""
import json
import os

import pytest
import torch

import golden_util as gu
from mog_b200 import synth
from oracle import stage1_oracle as S
from oracle.attngan_oracle import leafify

HERE = os.path.dirname(os.path.abspath(__file__))
FL = {"mnist": S.MNIST, "clevr": S.CLEVR}


def _keys(prog):
    return json.load(open(os.path.join(HERE, "golden", "stage1_%s_keys.json" % prog)))


def _set_cfg(prog, c):
    if prog == "mnist":
        from mog_b200.multi_mnist import model as M
        from mog_b200.multi_mnist.miscc import utils as U
        from mog_b200.multi_mnist.miscc.config import cfg, reset_cfg
    else:
        from mog_b200.clevr import model as M
        from mog_b200.clevr.miscc import utils as U
        from mog_b200.clevr.miscc.config import cfg, reset_cfg
    reset_cfg()
    cfg.GAN.GF_DIM, cfg.GAN.DF_DIM, cfg.Z_DIM, cfg.GAN.CONDITION_DIM = c["GF_DIM"], c["DF_DIM"], c["Z_DIM"], c["CONDITION_DIM"]
    return M, U


@pytest.mark.parametrize("prog", ["mnist", "clevr"])
def test_oracle_matches_reference(prog):
    G, meta = gu.load("stage1_" + prog)
    c, seed, fl = meta["cfg"], meta["seed"], FL[prog]
    keys = _keys(prog)
    PG = leafify(synth.fill_state_dict({k: torch.empty(s) for k, s in keys["STAGE1_G"].items()}, seed + 1))
    PD = leafify(synth.fill_state_dict({k: torch.empty(s) for k, s in keys["STAGE1_D"].items()}, seed + 2))
    b = synth.stage1_batch(prog, c["B"], nz=c["Z_DIM"], seed=seed)
    fake = S.stage1_g(PG, fl, b["noise"], b["transf_matrices_inv"], b["label_one_hot"], c["GF_DIM"] * 8)
    gu.check(fake, G["fake"], 2e-5, "fake")
    errD = S.discriminator_loss(PD, fl, b["imgs"], fake, b["label_one_hot"], b["transf_matrices"], b["transf_matrices_inv"], c["DF_DIM"])
    gu.check(errD, G["errD"], 2e-5, "errD")
    names = [k for k, p in PD.items() if p.requires_grad]
    for k, g in zip(names, torch.autograd.grad(errD, [PD[k] for k in names])):
        gu.check(g, G["D/grad/" + k], 1e-4, "D grad " + k)
    errG = S.generator_loss(PD, fl, fake, b["label_one_hot"], b["transf_matrices"], b["transf_matrices_inv"], c["DF_DIM"])
    gu.check(errG, G["errG"], 2e-5, "errG")
    names = [k for k, p in PG.items() if p.requires_grad and ("G/grad/" + k) in G]
    for k, g in zip(names, torch.autograd.grad(errG, [PG[k] for k in names])):
        gu.check(g, G["G/grad/" + k], 2e-4, "G grad " + k)


@pytest.mark.parametrize("prog", ["mnist", "clevr"])
def test_state_dict_contract(prog):
    _, meta = gu.load("stage1_" + prog)
    M, _ = _set_cfg(prog, meta["cfg"])
    keys = _keys(prog)
    for cls, k in ((M.STAGE1_G, "STAGE1_G"), (M.STAGE1_D, "STAGE1_D")):
        sd = {a: list(b.shape) for a, b in cls().state_dict().items()}
        assert sd == keys[k] and list(sd) == list(keys[k]), (prog, k)


@pytest.mark.gpu
@pytest.mark.parametrize("prec", ["fp32", "bf16x3"])
@pytest.mark.parametrize("prog", ["mnist", "clevr"])
def test_libmog_matches_reference(prog, prec):
    from mog_b200 import ops
    G, meta = gu.load("stage1_" + prog)
    c, seed = meta["cfg"], meta["seed"]
    M, U = _set_cfg(prog, c)
    ops.set_precision(prec)
    try:
        netG, netD = M.STAGE1_G(), M.STAGE1_D()
        netG.load_state_dict(synth.fill_state_dict(netG.state_dict(), seed + 1))
        netD.load_state_dict(synth.fill_state_dict(netD.state_dict(), seed + 2))
        netG.cuda().train()
        netD.cuda().train()
        b = {k: v.cuda() for k, v in synth.stage1_batch(prog, c["B"], nz=c["Z_DIM"], seed=seed).items()}
        out = netG(b["noise"], b["transf_matrices_inv"], b["label_one_hot"])
        fake = out[1] if isinstance(out, tuple) else out
        # bf16x3 gradients: the bbox encoder applies LeakyReLU to a label layout that is exactly/nearly zero outside the
        # boxes, so a few sign flips of ~1e-8 pre-activations move small weight gradients by several percent
        tol_o, tol_g = (5e-5, 5e-4) if prec == "fp32" else (2e-4, 0.15)
        gu.check(fake, G["fake"], tol_o, "fake")
        ones, zeros = torch.ones(c["B"], device="cuda"), torch.zeros(c["B"], device="cuda")
        errD, _, _, _ = U.compute_discriminator_loss(netD, b["imgs"], fake, ones, zeros, b["label_one_hot"],
                                                     b["transf_matrices"], b["transf_matrices_inv"], [0])
        errD.backward(retain_graph=True)
        gu.check(errD, G["errD"], tol_o, "errD")
        for k, p in netD.named_parameters():
            gu.check(p.grad, G["D/grad/" + k], tol_g, "D grad " + k)
        for k, v in netD.state_dict().items():
            if "running" in k:
                gu.check(v, G["D/buf/" + k], tol_o, k)
        netG.zero_grad()
        for p in netD.parameters():
            p.requires_grad_(False)
        errG = U.compute_generator_loss(netD, fake, ones, b["label_one_hot"], b["transf_matrices"], b["transf_matrices_inv"], [0])
        errG.backward()
        gu.check(errG, G["errG"], tol_o, "errG")
        for k, p in netG.named_parameters():
            if ("G/grad/" + k) in G:
                gu.check(p.grad, G["G/grad/" + k], tol_g, "G grad " + k)
    finally:
        ops.set_precision("fp32")


@pytest.mark.gpu
@pytest.mark.parametrize("prog", ["mnist", "clevr"])
def test_gan_trainer_trains(prog, tmp_path):
    """``GANTrainer(output_dir).train(data_loader)`` of the Multi-MNIST / CLEVR programs (multi-mnist/trainer.py:74-190,
    clevr/trainer.py:73-186): loader tuples of the reference's datasets, fused Adam, checkpoint dict of ``save_model``."""
    import glob
    import torch.utils.data
    _, meta = gu.load("stage1_" + prog)
    c = meta["cfg"]
    M, U = _set_cfg(prog, c)
    if prog == "mnist":
        from mog_b200.multi_mnist.trainer import GANTrainer
        from mog_b200.multi_mnist.miscc.config import cfg
    else:
        from mog_b200.clevr.trainer import GANTrainer
        from mog_b200.clevr.miscc.config import cfg
    cfg.TRAIN.BATCH_SIZE, cfg.TRAIN.MAX_EPOCH, cfg.TRAIN.SNAPSHOT_INTERVAL = 4, 1, 1
    b = synth.stage1_batch(prog, 8, nz=c["Z_DIM"], seed=3)

    class DS(torch.utils.data.Dataset):
        def __len__(self):
            return 8

        def __getitem__(self, i):
            if prog == "mnist":    # (image, bbox, label)
                return b["imgs"][i], b["bbox"][i], b["label_one_hot"][i]
            return b["imgs"][i], [b["transf_matrices"][i], b["transf_matrices_inv"][i]], b["label_one_hot"][i], 0

    tr = GANTrainer(str(tmp_path))
    loader = torch.utils.data.DataLoader(DS(), batch_size=4, drop_last=True, shuffle=False)
    st = tr.train(loader)
    ck = sorted(glob.glob(str(tmp_path / "Model" / "*.pth")))
    assert ck
    sd = torch.load(ck[-1], map_location="cpu")
    assert set(sd) == {"epoch", "netG", "optimG", "netD", "optimD"}
    for k, v in st["netG"].state_dict().items():
        assert torch.equal(sd["netG"][k], v.cpu()), k
    for p in st["netG"].parameters():
        assert bool(torch.isfinite(p).all())
    assert all(float(o.state_dict()["state"][0]["step"]) == 2 for o in (st["optG"], st["optD"]))
