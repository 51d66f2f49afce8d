"""Attribute-dict config helpers shared by the per-program ``miscc/config.py`` modules (the reference
uses easydict + a strict YAML merge, e.g. multi-mnist/miscc/config.py:50-89)."""
from __future__ import annotations

import numpy as np


class edict(dict):
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


def merge_a_into_b(a, b):
    if not isinstance(a, dict):
        return
    for k, v in a.items():
        if k not in b:
            raise KeyError('{} is not a valid config key'.format(k))
        if isinstance(v, dict) and not isinstance(v, edict):
            v = edict(v)
        old_type = type(b[k])
        if old_type is not type(v):
            if isinstance(b[k], np.ndarray):
                v = np.array(v, dtype=b[k].dtype)
            elif isinstance(b[k], float) and isinstance(v, int):
                v = float(v)
            else:
                raise ValueError('Type mismatch ({} vs. {}) for config key: {}'.format(type(b[k]), type(v), k))
        if isinstance(v, edict):
            merge_a_into_b(v, b[k])
        else:
            b[k] = v


def make_cfg(defaults_fn):
    cfg = defaults_fn()

    def cfg_from_file(filename):
        import yaml
        with open(filename, 'r') as f:
            merge_a_into_b(edict(yaml.safe_load(f)), cfg)

    def reset_cfg():
        d = defaults_fn()
        cfg.clear()
        for k, v in d.items():
            cfg[k] = v

    return cfg, cfg_from_file, reset_cfg
