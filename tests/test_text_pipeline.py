"""Input side of the path: ``RNN_ENCODER`` and ``prepare_data`` against the executed reference
(tests/golden/make_golden_text.py).  CPU."""
import numpy as np
import torch
import torch.utils.data

import golden_util as gu
from mog_b200 import synth


def _fill_rnn(sd, seed):
    rng = np.random.RandomState(seed)
    return {k: torch.from_numpy((0.3 * rng.standard_normal(tuple(v.shape))).astype(np.float32)) for k, v in sorted(sd.items())}


def _cfg(meta, cuda=False):
    from mog_b200.attngan.miscc.config import cfg, reset_cfg
    reset_cfg()
    cfg.CUDA = cuda
    cfg.TEXT.WORDS_NUM = meta["T"]
    return cfg


def test_rnn_encoder_matches_reference():
    from mog_b200.attngan import model as M
    from mog_b200.attngan.datasets import prepare_data
    G, meta = gu.load("text_pipeline")
    _cfg(meta)
    enc = M.RNN_ENCODER(meta["ntoken"], nhidden=meta["nhidden"])
    sd = {k: list(v.shape) for k, v in enc.state_dict().items()}
    assert sd == meta["state_dict_keys"]      # checkpoint wire format (names + shapes)
    enc.load_state_dict(_fill_rnn(enc.state_dict(), meta["seed"]))
    enc.eval()
    ds = synth.SyntheticTextDataset(n=meta["B"], n_words=meta["ntoken"], words_num=meta["T"], seed=3)
    batch = next(iter(torch.utils.data.DataLoader(ds, batch_size=meta["B"], shuffle=False)))
    imgs, captions, cap_lens, class_ids, keys, tms, label = prepare_data(batch)
    words, sent = enc(captions, cap_lens, enc.init_hidden(meta["B"]))
    gu.check(words, G["rnn/words_emb"], 1e-6, "words_emb")
    gu.check(sent, G["rnn/sent_emb"], 1e-6, "sent_emb")
    assert words.shape[2] == int(cap_lens.max())      # T = longest caption of the batch


def test_prepare_data_matches_reference():
    from mog_b200.attngan.datasets import prepare_data
    G, meta = gu.load("text_pipeline")
    _cfg(meta)
    ds = synth.SyntheticTextDataset(n=meta["B"], n_words=meta["ntoken"], words_num=meta["T"], seed=3)
    batch = next(iter(torch.utils.data.DataLoader(ds, batch_size=meta["B"], shuffle=False)))
    imgs, captions, cap_lens, class_ids, keys, tms, label = prepare_data(batch)
    assert list(keys) == meta["keys"]
    assert bool((cap_lens[:-1] >= cap_lens[1:]).all())              # sorted by decreasing caption length
    gu.check(captions.float(), G["prep/captions"], 0.0, "captions")
    gu.check(cap_lens.float(), G["prep/cap_lens"], 0.0, "cap_lens")
    gu.check(torch.from_numpy(np.asarray(class_ids)).float(), G["prep/class_ids"], 0.0, "class_ids")
    gu.check(tms[0], G["prep/theta"], 0.0, "theta")
    gu.check(tms[1], G["prep/theta_inv"], 0.0, "theta_inv")
    gu.check(label, G["prep/label"], 0.0, "label")
    for i, im in enumerate(imgs):
        gu.check(im, G["prep/img%d" % i], 0.0, "img%d" % i)
    dse = synth.SyntheticTextDataset(n=meta["B"], n_words=meta["ntoken"], words_num=meta["T"], seed=3, eval=True)
    out = prepare_data(next(iter(torch.utils.data.DataLoader(dse, batch_size=meta["B"], shuffle=False))), eval=True)
    assert len(out) == 8
    gu.check(out[7], G["prep_eval/bbox"], 0.0, "bbox")


def test_train_without_encoders_is_loud():
    """``algo.train()`` must never silently drop the DAMSM terms -- no NET_E and no encoders raises."""
    import pytest
    from mog_b200.attngan.miscc.config import cfg, reset_cfg
    from mog_b200.attngan.trainer import condGANTrainer
    reset_cfg()
    cfg.CUDA = False
    cfg.TRAIN.NET_E = ''
    tr = condGANTrainer("", [], 10, {})
    with pytest.raises(RuntimeError, match="NET_E"):
        tr.build_models()
