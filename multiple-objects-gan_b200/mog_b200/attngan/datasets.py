"""Input side of the hot path: ``prepare_data`` of ``code/coco/attngan/datasets.py:28-68`` (same name, arguments and return
list), i.e. what turns one ``DataLoader`` batch into the tensors ``condGANTrainer.train`` feeds to the text encoder and the
G/D step.  ``TextDataset`` itself (COCO file parsing, PIL decoding, caption vocabularies: ``datasets.py:71-399``) is I/O
outside the path and is not re-implemented: any ``Dataset`` yielding the reference's tuples works, including the reference's
own ``TextDataset``.

Differences, none of them result-affecting: host->device copies are issued ``non_blocking`` from pinned batches, the images
are uploaded once in their NCHW loader layout (the first libmog op converts to NHWC on the device)."""
from __future__ import annotations

import torch

from .miscc.config import cfg


def prepare_data(data, eval=False):
    if eval:
        imgs, captions, captions_lens, class_ids, keys, transformation_matrices, label, bbox = data
    else:
        imgs, captions, captions_lens, class_ids, keys, transformation_matrices, label = data
    # sort data by the caption length in decreasing order (pack_padded_sequence needs it, datasets.py:34-36)
    sorted_cap_lens, sorted_cap_indices = torch.sort(captions_lens, 0, True)
    dev = torch.device('cuda', torch.cuda.current_device()) if cfg.CUDA else torch.device('cpu')

    def up(t):
        return t.to(dev, non_blocking=True) if cfg.CUDA else t

    imgs = list(imgs)
    real_imgs = [up(imgs[i][sorted_cap_indices]) for i in range(len(imgs))]
    captions = captions[sorted_cap_indices].squeeze()
    if captions.dim() == 1:           # batch of one: squeeze() also dropped the batch axis
        captions = captions.unsqueeze(0)
    class_ids = class_ids[sorted_cap_indices].numpy()
    transformation_matrices = list(transformation_matrices)
    transformation_matrices[0] = up(transformation_matrices[0][sorted_cap_indices])
    transformation_matrices[1] = up(transformation_matrices[1][sorted_cap_indices])
    label = up(label[sorted_cap_indices])
    keys = [keys[i] for i in sorted_cap_indices.numpy()]
    captions = up(captions)
    sorted_cap_lens = up(sorted_cap_lens)
    if eval:
        bbox = bbox[sorted_cap_indices]
        return [real_imgs, captions, sorted_cap_lens, class_ids, keys, transformation_matrices, label, bbox]
    return [real_imgs, captions, sorted_cap_lens, class_ids, keys, transformation_matrices, label]
