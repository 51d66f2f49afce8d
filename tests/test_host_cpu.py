"""CPU-side checks (no GPU, no compute calls): the C-ABI library loads and exports every symbol
declared in include/mog.h, the ctypes table covers them, argument validation works without a
device, and the host-side mirror of the reference keeps its cfg / state_dict contract."""
import ctypes
import json
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "mog.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mog_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as g
    from mog_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        g.build()
    return _lib


def test_library_exports_every_declared_symbol(built):
    L = ctypes.CDLL(built.LIB_PATH)
    syms = _header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(L, s), "libmog.so does not export %s" % s
    assert set(syms) == set(built.SIGNATURES), set(syms) ^ set(built.SIGNATURES)


def test_argument_validation_without_device(built):
    L = built.lib()
    assert L.mog_version() == 100
    assert L.mog_nchw_to_nhwc(None, None, 1, 1, 1, 1, None) == -1
    assert b"mog_nchw_to_nhwc" in L.mog_last_error()
    d = built.MogConvDesc(2, 8, 8, 16, 32, 3, 3, 1, 1, 1, 0, 0)
    ho, wo = ctypes.c_int(), ctypes.c_int()
    assert L.mog_conv_out_hw(ctypes.byref(d), ctypes.byref(ho), ctypes.byref(wo)) == 0
    assert (ho.value, wo.value) == (16, 16)  # fused nearest x2
    d2 = built.MogConvDesc(2, 16, 16, 84, 192, 4, 4, 1, 1, 0, 0, 0)
    L.mog_conv_out_hw(ctypes.byref(d2), ctypes.byref(ho), ctypes.byref(wo))
    assert (ho.value, wo.value) == (15, 15)  # D_NET64.local quirk (model.py:677)
    assert L.mog_conv_workspace_bytes(ctypes.byref(d), 1) == 2 * 16 * 16 * 16 * 4


def test_ops_refuse_cpu_tensors(built):
    from mog_b200 import ops
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.conv2d(torch.zeros(1, 4, 4, 8), torch.zeros(8, 8, 3, 3), None, 1, 1)
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.activation(torch.zeros(2, 4), 3)


def test_missing_library_is_loud(monkeypatch, built):
    monkeypatch.setattr(built, "_lib", None)
    monkeypatch.setattr(built, "LIB_PATH", "/nonexistent/libmog.so")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        built.lib()


def test_state_dict_contract_matches_reference():
    from mog_b200.attngan import model as M
    from mog_b200.attngan.miscc.config import cfg, reset_cfg
    keys = json.load(open(os.path.join(ROOT, "tests", "golden", "attngan_state_dict_keys.json")))
    for name, c in (("config5", (48, 96, 100, 3, 256)), ("tiny", (8, 8, 20, 2, 32))):
        reset_cfg()
        cfg.GAN.GF_DIM, cfg.GAN.DF_DIM, cfg.GAN.Z_DIM, cfg.GAN.R_NUM, cfg.TEXT.EMBEDDING_DIM = c
        for cls, k in ((M.G_NET, "G_NET"), (M.D_NET64, "D_NET64"), (M.D_NET128, "D_NET128"), (M.D_NET256, "D_NET256")):
            sd = {a: list(b.shape) for a, b in cls().state_dict().items()}
            assert sd == keys[name][k], (name, k)
            assert list(sd) == list(keys[name][k]), "parameter order differs: %s %s" % (name, k)
    reset_cfg()


def test_reference_weights_init_applies():
    """The reference initialises by class name ('Conv', 'BatchNorm', 'Linear'), miscc/utils.py:321-331."""
    from mog_b200.attngan import model as M
    from mog_b200.attngan.miscc.config import cfg, reset_cfg
    from mog_b200.attngan.miscc.utils import weights_init
    reset_cfg()
    cfg.GAN.GF_DIM, cfg.GAN.DF_DIM, cfg.GAN.R_NUM, cfg.TEXT.EMBEDDING_DIM = 4, 4, 1, 16
    torch.manual_seed(0)
    d = M.D_NET64()
    d.apply(weights_init)
    w = d.conv2.weight.detach().reshape(d.conv2.weight.shape[0], -1)
    eye = w @ w.t()
    assert torch.allclose(eye, torch.eye(eye.shape[0]), atol=1e-4)  # orthogonal rows
    assert float(d.bn2.bias.abs().max()) == 0.0
    reset_cfg()


def test_cfg_from_file(tmp_path):
    from mog_b200.attngan.miscc.config import cfg, cfg_from_file, reset_cfg
    reset_cfg()
    p = tmp_path / "c.yml"
    p.write_text("GPU_ID: '0,1'\nTRAIN:\n    BATCH_SIZE: 32\n    SMOOTH:\n        GAMMA1: 4.0\nGAN:\n    DF_DIM: 96\n    GF_DIM: 48\n    R_NUM: 3\n")
    cfg_from_file(str(p))
    assert cfg.TRAIN.BATCH_SIZE == 32 and cfg.GAN.DF_DIM == 96 and cfg.GPU_ID == '0,1' and cfg.TRAIN.SMOOTH.GAMMA1 == 4.0
    p.write_text("NOT_A_KEY: 1\n")
    with pytest.raises(KeyError):
        cfg_from_file(str(p))
    reset_cfg()


def test_bbox_to_theta():
    import numpy as np
    from mog_b200.attngan.miscc.utils import compute_transformation_matrix, compute_transformation_matrix_inverse
    from mog_b200 import synth
    bbox = np.array([[0.1, 0.2, 0.5, 0.4], [-1, -1, -1, -1]], np.float32)
    t = compute_transformation_matrix(torch.from_numpy(bbox))
    ti = compute_transformation_matrix_inverse(torch.from_numpy(bbox))
    assert np.allclose(t.numpy(), synth.transformation_matrix(bbox))
    assert np.allclose(ti.numpy(), synth.transformation_matrix_inverse(bbox))
    assert ti[1].tolist() == [[-1.0, 0.0, -4.0], [0.0, -1.0, -4.0]]  # empty slot => fully out of range


def test_pack_plan_is_consistent_with_packed_bytes(built):
    """mog_pack_plan is host-only planning: the entries it emits for (descriptor, direction) must tile exactly the buffer
    mog_packed_weight_bytes sizes -- hi plane then lo plane per problem, each problem 256-byte aligned -- and carry sane
    tile grids (the multi-tensor repack kernel trusts them)."""
    import ctypes as C
    L = built.lib()
    cases = [  # N, H, W, Cin, Cout, KH, KW, stride, pad, up2x
        (4, 16, 16, 96, 96, 3, 3, 1, 1, 1),      # upBlock: four sub-pixel phases
        (4, 16, 16, 96, 192, 4, 4, 2, 1, 0),     # downBlock: stride phases / parity views
        (4, 8, 8, 384, 768, 4, 4, 2, 1, 0),
        (2, 4, 4, 1024, 768, 3, 3, 1, 1, 0),     # jointConv
        (6, 1, 1, 248, 512, 1, 1, 1, 0, 0),      # Linear
    ]
    buf = (built.MogPackEntry * 16)()
    base = 0x10000
    for prec in (built.PREC_BF16X3, built.PREC_BF16):
        planes = 2 if prec == built.PREC_BF16X3 else 1
        for c in cases:
            d = built.MogConvDesc(*c, 0, prec, 0)
            for which in (0, 1):
                n = L.mog_pack_plan(C.byref(d), which, C.c_void_p(0x1000), C.c_void_p(base), buf, 16)
                assert n >= 1, (c, which, L.mog_last_error())
                end = base
                for i in range(n):
                    e = buf[i]
                    assert e.hi == end, (c, which, i)
                    plane = e.Npad * e.Kpad * 2
                    assert (e.lo == e.hi + plane) if planes == 2 else (e.lo is None or e.lo == 0)
                    assert e.K == e.ntaps * e.Cs and e.Kpad >= e.K and e.Kpad % 64 == 0 and e.Cs % 8 == 0
                    assert e.nxb == (e.Cs + 31) // 32 and e.nyb == (e.Npad + 15) // 16 and 1 <= e.ntaps <= 16
                    assert all(0 <= e.taps[t][0] < e.KHW for t in range(e.ntaps))
                    end += (plane * planes + 255) // 256 * 256
                assert end - base == L.mog_packed_weight_bytes(C.byref(d), which), (c, which)
    # filters beyond 16 taps are left to mog_pack_weight
    d = built.MogConvDesc(2, 35, 35, 48, 64, 5, 5, 1, 2, 0, 0, built.PREC_BF16X3, 0)
    assert L.mog_pack_plan(C.byref(d), 0, C.c_void_p(0x1000), C.c_void_p(base), buf, 16) < 0


def test_reference_checkpoint_loads(tmp_path):
    """f4: a checkpoint written by the UNMODIFIED reference trainer's ``save_model`` (trainer.py:173-199, run here through
    baseline/write_ref_checkpoint.py) resumes in the libmog trainer: networks, epoch and Adam state."""
    import subprocess
    import sys
    import torch
    from baseline import ref_harness as H
    if not H.available():
        pytest.skip("reference sources not staged (baseline/_ref) on this box")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    model_dir = tmp_path / "Model"
    os.makedirs(model_dir)
    r = subprocess.run([sys.executable, os.path.join(root, "baseline", "write_ref_checkpoint.py"), str(model_dir), "7"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "written" in r.stdout, r.stderr[-2000:]
    ck = torch.load(str(model_dir / "checkpoint_0007.pth"), map_location="cpu")
    from mog_b200.attngan.miscc.config import cfg, reset_cfg
    from mog_b200.attngan.trainer import condGANTrainer
    reset_cfg()
    cfg.CUDA = False
    cfg.GAN.GF_DIM, cfg.GAN.DF_DIM, cfg.GAN.Z_DIM, cfg.GAN.R_NUM = 4, 4, 20, 1
    cfg.TEXT.EMBEDDING_DIM, cfg.TEXT.WORDS_NUM = 16, 6
    cfg.TRAIN.FLAG = True
    tr = condGANTrainer(str(tmp_path), [], 10, {}, resume=True)
    _, _, netG, netsD, epoch = tr.build_models(load_encoders=False)
    assert epoch == 8
    for k, v in netG.state_dict().items():
        assert torch.equal(v, ck["netG"][k]), k
    for i, d in enumerate(netsD):
        for k, v in d.state_dict().items():
            assert torch.equal(v, ck["netD"][i][k]), (i, k)
    optG, optDs = tr.define_optimizers(netG, netsD)
    sdG = optG.state_dict()
    assert len(sdG["state"]) == len(ck["optimG"]["state"]) > 0
    for i, s in sdG["state"].items():
        assert float(s["step"]) == 1.0
        assert torch.equal(s["exp_avg"], ck["optimG"]["state"][i]["exp_avg"])
