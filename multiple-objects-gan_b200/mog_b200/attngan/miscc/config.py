"""Global ``cfg`` of the AttnGAN program: same keys, defaults and YAML merge semantics as the
reference's ``code/coco/attngan/miscc/config.py:9-106`` (re-implemented for Python 3, without
the ``easydict`` dependency)."""
from __future__ import annotations

import numpy as np


class edict(dict):
    """Attribute-access dict (nested dicts are wrapped)."""

    def __init__(self, d=None, **kw):
        super().__init__()
        for k, v in dict(d or {}, **kw).items():
            self[k] = v

    def __setitem__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, edict):
            v = edict(v)
        super().__setitem__(k, v)

    __setattr__ = __setitem__

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)


def _defaults():
    c = edict()
    c.DATASET_NAME = 'birds'
    c.CONFIG_NAME = ''
    c.DATA_DIR = ''
    c.IMG_DIR = ''
    c.GPU_ID = '0'
    c.CUDA = True
    c.WORKERS = 6
    c.RNN_TYPE = 'LSTM'
    c.B_VALIDATION = False
    c.TREE = edict(BRANCH_NUM=3, BASE_SIZE=64)
    c.TRAIN = edict(BATCH_SIZE=64, MAX_EPOCH=600, SNAPSHOT_INTERVAL=2000, DISCRIMINATOR_LR=2e-4,
                    GENERATOR_LR=2e-4, ENCODER_LR=2e-4, RNN_GRAD_CLIP=0.25, FLAG=True, NET_E='', NET_G='',
                    B_NET_D=True,
                    SMOOTH=edict(GAMMA1=5.0, GAMMA3=10.0, GAMMA2=5.0, LAMBDA=1.0))
    c.GAN = edict(DF_DIM=64, GF_DIM=128, Z_DIM=100, CONDITION_DIM=100, R_NUM=2, B_ATTENTION=True, B_DCGAN=False)
    c.TEXT = edict(CAPTIONS_PER_IMAGE=10, EMBEDDING_DIM=256, WORDS_NUM=18)
    # libmog extensions (not in the reference): conv operand precision and STN convention
    c.MOG = edict(PRECISION='bf16x3', ALIGN_CORNERS=False, MASK_QUIRK=True, CUDA_GRAPH=True, STREAMS=True)
    return c


cfg = _defaults()
__C = cfg


def _merge_a_into_b(a, b):
    """Strict merge: keys of ``a`` must exist in ``b`` with the same type (config.py:67-97)."""
    if not isinstance(a, dict):
        return
    for k, v in a.items():
        if k not in b:
            raise KeyError('{} is not a valid config key'.format(k))
        old_type = type(b[k])
        if isinstance(v, dict) and not isinstance(v, edict):
            v = edict(v)
        if old_type is not type(v):
            if isinstance(b[k], np.ndarray):
                v = np.array(v, dtype=b[k].dtype)
            elif isinstance(b[k], float) and isinstance(v, int):
                v = float(v)
            else:
                raise ValueError('Type mismatch ({} vs. {}) for config key: {}'.format(type(b[k]), type(v), k))
        if isinstance(v, edict):
            try:
                _merge_a_into_b(v, b[k])
            except Exception:
                print('Error under config key: {}'.format(k))
                raise
        else:
            b[k] = v


def cfg_from_file(filename):
    """Load a YAML config file and merge it into the defaults (config.py:100-106)."""
    import yaml
    with open(filename, 'r') as f:
        yaml_cfg = edict(yaml.safe_load(f))
    _merge_a_into_b(yaml_cfg, cfg)


def reset_cfg():
    d = _defaults()
    cfg.clear()
    for k, v in d.items():
        cfg[k] = v
