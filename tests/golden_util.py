"""Compact tensor summaries for the committed golden fixtures (full tensor when small, otherwise
a fixed pseudo-random subsample plus norms) and the comparison used by the tests."""
import numpy as np
import torch

FULL_LIMIT = 8192
N_SAMPLE = 2048


def _idx(numel):
    return np.random.RandomState(numel % 100003).choice(numel, N_SAMPLE, replace=False)


def summarize(t):
    a = t.detach().cpu().double().numpy().reshape(-1)
    d = {"shape": list(t.shape), "l2": float(np.sqrt((a * a).sum())), "sum": float(a.sum())}
    if a.size <= FULL_LIMIT:
        d["full"] = a.astype(np.float32)
    else:
        d["sample"] = a[_idx(a.size)].astype(np.float32)
    return d


def rel_l2(a, b):
    a = np.asarray(a, np.float64).reshape(-1)
    b = np.asarray(b, np.float64).reshape(-1)
    den = np.sqrt((b * b).sum())
    return float(np.sqrt(((a - b) ** 2).sum()) / max(den, 1e-30))


def check(t, summ, tol, name=""):
    """rel-L2 of the stored entries + relative agreement of the global L2 norm."""
    assert list(t.shape) == list(summ["shape"]), (name, t.shape, summ["shape"])
    a = t.detach().cpu().double().numpy().reshape(-1)
    ref = summ["full"] if "full" in summ else summ["sample"]
    got = a if "full" in summ else a[_idx(a.size)]
    err = rel_l2(got, ref)
    l2 = float(np.sqrt((a * a).sum()))
    if summ["l2"] < 1e-20:
        assert l2 < 1e-12, (name, "expected zeros", l2)
        return 0.0
    assert err <= tol, "%s: rel-L2 %.3e > %.1e" % (name, err, tol)
    assert abs(l2 - summ["l2"]) <= 10 * tol * summ["l2"] + 1e-12, (name, l2, summ["l2"])
    return err


def load(name):
    import json
    import os
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    with open(os.path.join(here, name + ".json")) as f:
        meta = json.load(f)
    arr = np.load(os.path.join(here, name + ".npz"))
    out = {}
    for k, e in meta["index"].items():
        d = {"shape": e["shape"], "l2": e["l2"], "sum": e["sum"]}
        d[e["kind"]] = arr[k]
        out[k] = d
    return out, meta["meta"]


def full(summ, key):
    """A fully stored golden tensor as torch tensor."""
    s = summ[key]
    return torch.from_numpy(np.asarray(s["full"], np.float32).reshape(s["shape"]).copy())


def save(name, entries, meta):
    """Write tests/golden/<name>.npz (+ .json index) from a dict of summarize() results."""
    import json
    import os
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    arrays, index = {}, {}
    for k, s in entries.items():
        index[k] = {"shape": s["shape"], "l2": s["l2"], "sum": s["sum"], "kind": "full" if "full" in s else "sample"}
        arrays[k] = s["full"] if "full" in s else s["sample"]
    np.savez_compressed(os.path.join(here, name + ".npz"), **arrays)
    with open(os.path.join(here, name + ".json"), "w") as f:
        json.dump({"meta": meta, "index": index}, f, indent=1, sort_keys=True)
    print("wrote", name, "entries", len(entries), "bytes", os.path.getsize(os.path.join(here, name + ".npz")))


class Aggregate:
    """Norm-weighted rel-L2 over a group of tensors (e.g. all Adam moments of one network): tensors whose reference is
    tiny (cancelling gradients) cannot dominate the way they do in a per-tensor maximum."""

    def __init__(self):
        self.num = self.den = 0.0
        self.worst = (0.0, "")

    def add(self, t, summ, name=""):
        a = t.detach().cpu().double().numpy().reshape(-1)
        ref = np.asarray(summ["full"] if "full" in summ else summ["sample"], np.float64)
        got = a if "full" in summ else a[_idx(a.size)]
        scale = a.size / ref.size          # a sample stands for the whole tensor
        self.num += scale * float(((got - ref) ** 2).sum())
        self.den += scale * float((ref ** 2).sum())
        e = rel_l2(got, ref)
        if e > self.worst[0]:
            self.worst = (e, name)
        return e

    def rel(self):
        return float(np.sqrt(self.num / max(self.den, 1e-60)))
