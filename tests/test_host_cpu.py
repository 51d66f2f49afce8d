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
