"""Deterministic synthetic inputs and weights for the G/D hot path.

Everything is drawn from ``numpy.random.RandomState`` (bit-stable across numpy/torch versions and
machines), so the golden fixtures under ``tests/golden`` can be regenerated and re-checked anywhere.
Value ranges follow what the reference's data pipeline produces:

* captions sorted by decreasing length, T = longest (``attngan/datasets.py:35-36``),
* bounding boxes as (x, y, w, h) fractions with ``-1`` marking an empty slot
  (``attngan/datasets.py:107-109``), labels 0..79 and 80 for "no object" one-hot over 81,
* theta / theta^-1 as in ``attngan/miscc/utils.py:16-49``.
"""
from __future__ import annotations

import numpy as np
import torch

MAX_OBJECTS = 3
N_LABELS = 81


def transformation_matrix(bbox: np.ndarray) -> np.ndarray:
    """bbox (N,4) x,y,w,h -> theta (N,2,3): crops the box out of the image.
    Follows ``attngan/miscc/utils.py:34-49`` (fp32 arithmetic)."""
    bbox = bbox.astype(np.float32)
    x, y, w, h = bbox[:, 0], bbox[:, 1], bbox[:, 2], bbox[:, 3]
    th = np.zeros((bbox.shape[0], 2, 3), np.float32)
    th[:, 0, 0] = w
    th[:, 0, 2] = np.float32(2) * ((x + np.float32(0.5) * w) - np.float32(0.5))
    th[:, 1, 1] = h
    th[:, 1, 2] = np.float32(2) * ((y + np.float32(0.5) * h) - np.float32(0.5))
    return th


def transformation_matrix_inverse(bbox: np.ndarray) -> np.ndarray:
    """bbox (N,4) -> theta^-1 (N,2,3): places a full canvas into the box.
    Follows ``attngan/miscc/utils.py:16-31``; an empty slot (-1,-1,-1,-1) yields
    [[-1,0,-4],[0,-1,-4]] whose sampling grid lies fully outside the image => exact zeros."""
    bbox = bbox.astype(np.float32)
    x, y, w, h = bbox[:, 0], bbox[:, 1], bbox[:, 2], bbox[:, 3]
    sx = np.float32(1.0) / w
    sy = np.float32(1.0) / h
    th = np.zeros((bbox.shape[0], 2, 3), np.float32)
    th[:, 0, 0] = sx
    th[:, 0, 2] = np.float32(2) * sx * (np.float32(0.5) - (x + np.float32(0.5) * w))
    th[:, 1, 1] = sy
    th[:, 1, 2] = np.float32(2) * sy * (np.float32(0.5) - (y + np.float32(0.5) * h))
    return th


def bboxes_and_labels(rng: np.random.RandomState, B: int, max_objects: int = MAX_OBJECTS,
                      n_labels: int = N_LABELS, always_valid: bool = False):
    """Random layout: per sample k in {0..max_objects} valid boxes (first k slots), rest empty."""
    bbox = -np.ones((B, max_objects, 4), np.float32)
    label = np.full((B, max_objects), n_labels - 1, np.int64)
    for b in range(B):
        k = max_objects if always_valid else int(rng.randint(0, max_objects + 1))
        for s in range(k):
            w, h = rng.uniform(0.15, 0.6, size=2)
            x = rng.uniform(0.0, 0.999 - w)
            y = rng.uniform(0.0, 0.999 - h)
            bbox[b, s] = (x, y, w, h)
            label[b, s] = int(rng.randint(0, n_labels - 1))
    onehot = np.zeros((B, max_objects, n_labels), np.float32)
    for b in range(B):
        for s in range(max_objects):
            onehot[b, s, label[b, s]] = 1.0
    flat = bbox.reshape(-1, 4)
    theta = transformation_matrix(flat).reshape(B, max_objects, 2, 3)
    theta_inv = transformation_matrix_inverse(flat).reshape(B, max_objects, 2, 3)
    return bbox, label, onehot, theta, theta_inv


def attngan_batch(B: int, T: int = 18, nef: int = 256, nz: int = 100, seed: int = 1234,
                  sizes=(64, 128, 256)) -> dict:
    """One synthetic AttnGAN training batch (config 5 of BASELINE.json when called with defaults
    and B=32). Returns CPU torch tensors."""
    rng = np.random.RandomState(seed)
    out = {}
    out["noise"] = rng.standard_normal((B, nz)).astype(np.float32)
    out["sent_emb"] = np.tanh(rng.standard_normal((B, nef))).astype(np.float32)
    out["words_embs"] = np.tanh(rng.standard_normal((B, nef, T))).astype(np.float32)
    lo = min(5, T)
    lens = np.sort(rng.randint(lo, T + 1, size=B))[::-1].copy()
    lens[0] = T
    out["cap_lens"] = lens.astype(np.int64)
    out["mask"] = (np.arange(T)[None, :] >= lens[:, None])
    out["imgs"] = [rng.uniform(-1, 1, size=(B, 3, s, s)).astype(np.float32) for s in sizes]
    bbox, label, onehot, theta, theta_inv = bboxes_and_labels(rng, B)
    out["bbox"], out["label"], out["label_one_hot"] = bbox, label, onehot
    out["transf_matrices"], out["transf_matrices_inv"] = theta, theta_inv
    out["eps"] = rng.standard_normal((B, 100)).astype(np.float32)  # CA_NET reparametrisation draw
    out["class_ids"] = np.arange(B)
    res = {}
    for k, v in out.items():
        if isinstance(v, list):
            res[k] = [torch.from_numpy(a) for a in v]
        elif k == "class_ids":
            res[k] = v
        else:
            res[k] = torch.from_numpy(np.ascontiguousarray(v))
    return res


def fill_state_dict(sd: dict, seed: int) -> dict:
    """Deterministic, version-independent weights for a ``state_dict`` (keys/shapes from the
    reference classes). Conv/Linear ~ N(0, 1/fan_in) so activations stay O(1) through the depth
    (comparable to the orthogonal init of ``attngan/miscc/utils.py:321-331``); BatchNorm gamma ~
    N(1, 0.02) as in the reference, beta/bias small but non-zero so bias paths are exercised."""
    rng = np.random.RandomState(seed)
    out = {}
    for k in sorted(sd.keys()):
        v = sd[k]
        shape = tuple(v.shape)
        if k.endswith("num_batches_tracked"):
            out[k] = torch.zeros(shape, dtype=torch.long)
        elif k.endswith("running_mean"):
            out[k] = torch.zeros(shape, dtype=torch.float32)
        elif k.endswith("running_var"):
            out[k] = torch.ones(shape, dtype=torch.float32)
        elif k.endswith("weight") and len(shape) >= 2:
            fan_in = int(np.prod(shape[1:]))
            out[k] = torch.from_numpy((rng.standard_normal(shape) / np.sqrt(fan_in)).astype(np.float32))
        elif k.endswith("weight"):
            out[k] = torch.from_numpy((1.0 + 0.02 * rng.standard_normal(shape)).astype(np.float32))
        elif k.endswith("bias"):
            out[k] = torch.from_numpy((0.02 * rng.standard_normal(shape)).astype(np.float32))
        else:
            raise KeyError("unexpected state_dict entry %s" % k)
    return out


def soften_logits(sd: dict, factor: float = 0.02) -> dict:
    """Scale the ``outlogits`` weights of a filled discriminator ``state_dict`` in place.  With N(0, 1/fan_in) weights
    the 4x4 logit convs of the tiny test nets saturate the sigmoid (BCE at its -100 clamp, vanishing and noise-dominated
    gradients); multi-step fixtures use this to keep the losses in the regime training runs in."""
    for k, v in sd.items():
        if "outlogits" in k and k.endswith("weight"):
            v.mul_(factor)
    return sd


class StandInEncoder:
    """A tiny fixed, differentiable stand-in for the frozen DAMSM image encoder (Inception-v3,
    ``attngan/model.py:207-313``), used so the DAMSM branch of ``generator_loss``
    (``attngan/miscc/losses.py:205-224``) can be exercised without the 22.5 M-parameter network
    (SURVEY.md section 8(f) row f1 -- the real encoder is a later scope row).
    ``img (B,3,S,S) -> (region_features (B,nef,17,17), cnn_code (B,nef))``."""

    def __init__(self, nef: int, seed: int = 7, device="cpu"):
        rng = np.random.RandomState(seed)
        self.w_feat = torch.from_numpy(rng.standard_normal((nef, 3, 1, 1)).astype(np.float32)).to(device)
        self.w_code = torch.from_numpy(rng.standard_normal((nef, 3 * 16)).astype(np.float32) * 0.25).to(device)

    def __call__(self, img):
        import torch.nn.functional as F
        feat = F.conv2d(F.adaptive_avg_pool2d(img, 17), self.w_feat)
        code = F.linear(F.adaptive_avg_pool2d(img, 4).reshape(img.shape[0], -1), self.w_code)
        return feat, code


def stage1_batch(program: str, B: int, nz: int = 100, seed: int = 1234) -> dict:
    """Synthetic batch for the single-stage programs (SURVEY.md section 8(d) "other configs").
    multi-mnist: images U(-1,1) Bx1x64x64, 3 objects always valid, w in [10,20)/64, h in [16,20)/64
    (multi-mnist/trainer.py:236-237), labels one-hot over 10.
    clevr: Bx3x64x64, 4 slots with 2-4 valid objects, label = one-hot4(shape) + one-hot9(colour); an empty
    slot has bbox -1 and an all -1 label row (clevr/miscc/datasets.py:54-80, cf. the clamp at
    clevr/miscc/utils.py:99)."""
    rng = np.random.RandomState(seed)
    mn = program in ("mnist", "multi-mnist", "multi_mnist")
    S, L, C = (3, 10, 1) if mn else (4, 13, 3)
    out = {"noise": rng.standard_normal((B, nz)).astype(np.float32),
           "imgs": rng.uniform(-1, 1, size=(B, C, 64, 64)).astype(np.float32)}
    bbox = -np.ones((B, S, 4), np.float32)
    label = np.zeros((B, S, L), np.float32)
    for b in range(B):
        k = S if mn else int(rng.randint(2, S + 1))
        for s in range(S):
            if s < k:
                if mn:
                    w, h = rng.randint(10, 20) / 64.0, rng.randint(16, 20) / 64.0
                else:
                    w, h = rng.uniform(0.15, 0.5, size=2)
                x, y = rng.uniform(0.0, 0.999 - w), rng.uniform(0.0, 0.999 - h)
                bbox[b, s] = (x, y, w, h)
                if mn:
                    label[b, s, rng.randint(0, 10)] = 1.0
                else:
                    label[b, s, rng.randint(0, 4)] = 1.0
                    label[b, s, 4 + rng.randint(0, 9)] = 1.0
            else:
                label[b, s, :] = -1.0
    flat = bbox.reshape(-1, 4)
    out["bbox"] = bbox
    out["label_one_hot"] = label
    out["transf_matrices"] = transformation_matrix(flat).reshape(B, S, 2, 3)
    out["transf_matrices_inv"] = transformation_matrix_inverse(flat).reshape(B, S, 2, 3)
    return {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in out.items()}


def stackgan_batch(B: int, stage: int = 1, t_dim: int = 1024, nz: int = 100, c_dim: int = 128, seed: int = 1234) -> dict:
    """Synthetic batch for the COCO StackGAN program: ``txt_embedding``
    N(0,1) B x t_dim (char-CNN-RNN stand-in), real images U(-1,1) at 64^2 (stage I) or 256^2 (stage II),
    up to 3 boxes per image with labels 0..79 (80 = empty slot).  Stage II carries two box sets like
    ``stackgan/miscc/datasets.py:134-179``: the same objects scaled for the 64^2 stage-I generator
    (76 -> 64 crop) and for the stage-II image; ``eps1`` / ``eps2`` are the CA_NET draws (stage-I net
    first, ``stackgan/model.py:379,386``)."""
    rng = np.random.RandomState(seed)
    out = {"noise": rng.standard_normal((B, nz)).astype(np.float32),
           "txt_embedding": rng.standard_normal((B, t_dim)).astype(np.float32)}
    size = 64 if stage == 1 else 256
    out["imgs"] = rng.uniform(-1, 1, size=(B, 3, size, size)).astype(np.float32)
    bbox, label, onehot, theta, theta_inv = bboxes_and_labels(rng, B)
    out["label_one_hot"] = onehot
    if stage == 1:
        out["transf_matrices"], out["transf_matrices_inv"] = theta, theta_inv
    else:
        # stage-I view of the same boxes: crop offset + 76/64 rescale, clipped like datasets.py:140-150
        b1 = bbox.copy()
        off = rng.uniform(0.0, 12.0 / 64.0, size=(B, 1, 2)).astype(np.float32)
        valid = bbox[:, :, 0] >= 0
        x = np.maximum(bbox[:, :, 0] * (76.0 / 64.0) - off[:, :, 0], 0.0)
        y = np.maximum(bbox[:, :, 1] * (76.0 / 64.0) - off[:, :, 1], 0.0)
        w = np.minimum(bbox[:, :, 2] * (76.0 / 64.0), 1.0)
        h = np.minimum(bbox[:, :, 3] * (76.0 / 64.0), 1.0)
        w = np.where(x + w > 0.999, 1.0 - x - 0.001, w)
        h = np.where(y + h > 0.999, 1.0 - y - 0.001, h)
        for i, v in enumerate((x, y, w, h)):
            b1[:, :, i] = np.where(valid, v, -1.0)
        flat1 = b1.reshape(-1, 4).astype(np.float32)
        out["transf_matrices_inv"] = transformation_matrix_inverse(flat1).reshape(B, MAX_OBJECTS, 2, 3)
        out["transf_matrices_s2"], out["transf_matrices_inv_s2"] = theta, theta_inv
    out["eps1"] = rng.standard_normal((B, c_dim)).astype(np.float32)
    out["eps2"] = rng.standard_normal((B, c_dim)).astype(np.float32)
    return {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in out.items()}


def fill_encoder_state_dict(sd: dict, seed: int) -> dict:
    """Deterministic stand-in for the ImageNet weights of the DAMSM image encoder (``CNN_ENCODER`` /
    torchvision ``inception_v3`` keys): He-scaled conv/linear weights so activations survive the ReLU depth,
    and NON-trivial BatchNorm affine + running statistics so the eval-mode folding is exercised."""
    rng = np.random.RandomState(seed)
    out = {}
    for k in sorted(sd.keys()):
        shape = tuple(sd[k].shape)
        if k.endswith("num_batches_tracked"):
            out[k] = torch.zeros(shape, dtype=torch.long)
        elif k.endswith("running_mean"):
            out[k] = torch.from_numpy((0.1 * rng.standard_normal(shape)).astype(np.float32))
        elif k.endswith("running_var"):
            out[k] = torch.from_numpy(rng.uniform(0.5, 1.5, size=shape).astype(np.float32))
        elif k.endswith("weight") and len(shape) >= 2:
            fan_in = int(np.prod(shape[1:]))
            out[k] = torch.from_numpy((rng.standard_normal(shape) * np.sqrt(2.0 / fan_in)).astype(np.float32))
        elif k.endswith("weight"):
            out[k] = torch.from_numpy(rng.uniform(0.5, 1.5, size=shape).astype(np.float32))
        elif k.endswith("bias"):
            out[k] = torch.from_numpy((0.1 * rng.standard_normal(shape)).astype(np.float32))
        else:
            raise KeyError("unexpected state_dict entry %s" % k)
    return out


def encoder_probe(B: int, nef: int, seed: int):
    """Input images U(-1,1) B x 3 x 256 x 256 and fixed projections that turn the encoder outputs into a scalar
    (so a gradient w.r.t. the image exists): loss = sum(features * pf) + sum(cnn_code * pc)."""
    rng = np.random.RandomState(seed)
    img = rng.uniform(-1, 1, size=(B, 3, 256, 256)).astype(np.float32)
    pf = rng.standard_normal((B, nef, 17, 17)).astype(np.float32)
    pc = rng.standard_normal((B, nef)).astype(np.float32)
    return torch.from_numpy(img), torch.from_numpy(pf), torch.from_numpy(pc)


class SyntheticTextDataset(torch.utils.data.Dataset):
    """Stand-in for the reference's ``TextDataset`` (``attngan/datasets.py:71-399``, COCO files + PIL decoding): yields the
    same per-sample tuple -- ``(imgs [3 x (3,S,S)], caption (WORDS_NUM,1) int64, cap_len, class_id, key,
    [theta (3,2,3), theta^-1 (3,2,3)], label one-hot (3,81)[, bbox (3,4)])`` -- from seeded random data, so that
    ``DataLoader`` -> ``prepare_data`` -> ``condGANTrainer.train`` can be driven end to end without the dataset files."""

    def __init__(self, n=8, n_words=40, words_num=18, seed=0, eval=False, sizes=(64, 128, 256)):
        self.n, self.n_words, self.words_num, self.seed, self.eval, self.sizes = n, n_words, words_num, seed, eval, sizes
        self.ixtoword = {i: ("w%d" % i) for i in range(1, n_words)}
        self.ixtoword[0] = "<end>"
        self.wordtoix = {v: k for k, v in self.ixtoword.items()}

    def __len__(self):
        return self.n

    def __getitem__(self, index):
        rng = np.random.RandomState(self.seed * 100003 + index)
        imgs = [torch.from_numpy(rng.uniform(-1, 1, size=(3, s, s)).astype(np.float32)) for s in self.sizes]
        ln = int(rng.randint(3, self.words_num + 1))
        cap = np.zeros((self.words_num, 1), np.int64)
        cap[:ln, 0] = rng.randint(1, self.n_words, size=ln)
        bbox, label, onehot, theta, theta_inv = bboxes_and_labels(rng, 1)
        out = (imgs, cap, ln, index, "img%04d" % index, [torch.from_numpy(theta[0]), torch.from_numpy(theta_inv[0])],
               torch.from_numpy(onehot[0]))
        if self.eval:
            out = out + (torch.from_numpy(bbox[0]),)
        return out
