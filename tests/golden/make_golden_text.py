#!/usr/bin/env python
"""Golden fixtures for the input side of the path (f3), by EXECUTING THE UNMODIFIED REFERENCE:

    python tests/golden/make_golden_text.py

* ``RNN_ENCODER`` (code/coco/attngan/model.py:120-204), eval mode, deterministic weights: words / sentence embeddings of a
  sorted caption batch;
* ``prepare_data`` (code/coco/attngan/datasets.py:28-68) on one collated batch of ``mog_b200.synth.SyntheticTextDataset``
  (training and eval form)."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
from baseline import ref_harness as H  # noqa: E402
from golden_util import save, summarize  # noqa: E402

NTOKEN, NHID, T, B, SEED = 40, 32, 18, 5, 300


def fill_rnn(sd, seed):
    rng = np.random.RandomState(seed)
    return {k: torch.from_numpy((0.3 * rng.standard_normal(tuple(v.shape))).astype(np.float32)) for k, v in sorted(sd.items())}


if __name__ == "__main__":
    ns = H.load("attngan", "cpu")
    ns.cfg.TEXT.WORDS_NUM, ns.cfg.RNN_TYPE = T, 'LSTM'
    import datasets as RD  # reference (nltk / PIL shims)
    import torch.utils.data
    from mog_b200 import synth
    E = {}
    enc = ns.model.RNN_ENCODER(NTOKEN, nhidden=NHID)
    enc.load_state_dict(fill_rnn(enc.state_dict(), SEED))
    enc.eval()
    ds = synth.SyntheticTextDataset(n=B, n_words=NTOKEN, words_num=T, seed=3)
    batch = next(iter(torch.utils.data.DataLoader(ds, batch_size=B, shuffle=False)))
    imgs, captions, cap_lens, class_ids, keys, tms, label = RD.prepare_data(batch)
    words, sent = enc(captions, cap_lens, enc.init_hidden(B))
    E["rnn/words_emb"], E["rnn/sent_emb"] = summarize(words), summarize(sent)
    E["prep/captions"], E["prep/cap_lens"] = summarize(captions.float()), summarize(cap_lens.float())
    E["prep/class_ids"] = summarize(torch.from_numpy(np.asarray(class_ids)).float())
    E["prep/theta"], E["prep/theta_inv"], E["prep/label"] = summarize(tms[0]), summarize(tms[1]), summarize(label)
    for i, im in enumerate(imgs):
        E["prep/img%d" % i] = summarize(im)
    dse = synth.SyntheticTextDataset(n=B, n_words=NTOKEN, words_num=T, seed=3, eval=True)
    out = RD.prepare_data(next(iter(torch.utils.data.DataLoader(dse, batch_size=B, shuffle=False))), eval=True)
    E["prep_eval/bbox"] = summarize(out[7])
    save("text_pipeline", E, {"ntoken": NTOKEN, "nhidden": NHID, "T": T, "B": B, "seed": SEED, "keys": list(keys),
                              "state_dict_keys": {k: list(v.shape) for k, v in enc.state_dict().items()},
                              "what": "reference RNN_ENCODER (eval) + prepare_data"})
