"""Helpers to read tests/golden/*.npz (outputs of the reference itself, see oracle/make_golden.py)."""
from __future__ import annotations

from pathlib import Path

import numpy as np

GOLDEN = Path(__file__).resolve().parent / "golden"


def load(name):
    return np.load(GOLDEN / name, allow_pickle=False)


def parse_lists(arr):
    """np array of 'a,b,c' strings -> list of int lists."""
    return [[int(float(v)) for v in s.split(",") if v != ""] for s in arr.tolist()]


class Trace:
    """Recorded event stream of one train_FedMLP call (prefix 's1', 's2_0', 's2_1')."""

    def __init__(self, z, prefix):
        self.z, self.prefix = z, prefix
        self.kinds = [str(k) for k in z[f"{prefix}/kinds"]]

    def get(self, i, field):
        kind = self.kinds[i]
        if kind == "loss":
            return float(self.z[f"{self.prefix}/{i}/loss"])
        key = f"{self.prefix}/{i}/{kind}/{field}"
        return self.z[key] if key in self.z.files else None

    def indices(self, kind):
        return [i for i, k in enumerate(self.kinds) if k == kind]


def stage1_steps(tr: Trace):
    """Training steps of the stage-1 round: each = 4 forwards (student v1, student v2, global v1,
    global v2), 2 bce calls (same labels), 1 loss.  Yields dicts of numpy arrays."""
    steps = []
    i, n = 0, len(tr.kinds)
    while i < n:
        if tr.kinds[i] == "forward" and i + 6 < n and tr.kinds[i:i + 7] == ["forward"] * 4 + ["bce"] * 2 + ["loss"]:
            f = [i, i + 1, i + 2, i + 3]
            steps.append(dict(
                z1=tr.get(f[0], "logits"), z2=tr.get(f[1], "logits"), z3=tr.get(f[2], "logits"), z4=tr.get(f[3], "logits"),
                dz1=tr.get(f[0], "dlogits"), dz2=tr.get(f[1], "dlogits"),
                y=tr.get(i + 4, "target"), bce1=tr.get(i + 4, "out"), bce2=tr.get(i + 5, "out"),
                p1=tr.get(i + 4, "p"), loss=tr.get(i + 6, None)))
            i += 7
        else:
            i += 1
    return steps


def eval_passes(tr: Trace, start, stop):
    """Consecutive (split_item*, forward) groups between event indices [start, stop): returns
    (dataset_idx [N], labels [N,C], features [N,D], logits [N,C]) concatenated in loader order."""
    idx, lab, feat, logit = [], [], [], []
    pend_idx, pend_lab = [], []
    for i in range(start, stop):
        k = tr.kinds[i]
        if k == "split_item":
            pend_idx.append(int(tr.get(i, "idx")))
            pend_lab.append(tr.get(i, "target"))
        elif k == "forward" and pend_idx:
            idx += pend_idx
            lab += pend_lab
            feat.append(tr.get(i, "feature"))
            logit.append(tr.get(i, "logits"))
            pend_idx, pend_lab = [], []
    return (np.array(idx, dtype=np.int64), np.stack(lab).astype(np.float32),
            np.concatenate(feat).astype(np.float32), np.concatenate(logit).astype(np.float32))


def stage1_proto_pass(tr: Trace):
    """The prototype / t pass at the end of the last stage-1 round: everything after the last loss."""
    last_loss = tr.indices("loss")[-1]
    return eval_passes(tr, last_loss + 1, len(tr.kinds))


def stage2_round(tr: Trace):
    """Split a stage-2 trace into its phases."""
    first_cos = tr.indices("cos")[0]
    extract = eval_passes(tr, 0, first_cos)
    cos = [dict(x1=tr.get(i, "x1"), x2=tr.get(i, "x2"), out=tr.get(i, "out")) for i in tr.indices("cos")]
    sel = []
    for i in tr.indices("max_m"):
        sel.append(dict(kind="max", values=tr.get(i, "values"), n=int(tr.get(i, "n")), out=tr.get(i, "out")))
    for j, i in enumerate(tr.indices("min_n")):
        sel.insert(2 * j + 1, dict(kind="min", values=tr.get(i, "values"), n=int(tr.get(i, "n")), out=tr.get(i, "out")))
    # training steps: pseudo_item* forward forward bce loss
    steps = []
    i, n = 0, len(tr.kinds)
    pend = []
    while i < n:
        k = tr.kinds[i]
        if k == "pseudo_item":
            pend.append(i)
            i += 1
        elif k == "forward" and pend and tr.kinds[i:i + 4] == ["forward", "forward", "bce", "loss"]:
            steps.append(dict(
                idx=np.array([int(tr.get(p, "idx")) for p in pend], dtype=np.int64),
                target=np.stack([tr.get(p, "target") for p in pend]).astype(np.float32),
                distill=np.stack([tr.get(p, "distill") for p in pend]).astype(np.float32),
                z=tr.get(i, "logits"), dz=tr.get(i, "dlogits"), zg=tr.get(i + 1, "logits"),
                y=tr.get(i + 2, "target"), bce=tr.get(i + 2, "out"), loss=tr.get(i + 3, None)))
            pend = []
            i += 4
        else:
            i += 1
    last_loss = tr.indices("loss")[-1]
    proto_pass = eval_passes(tr, last_loss + 1, len(tr.kinds))
    return dict(extract=extract, cos=cos, sel=sel, steps=steps, proto_pass=proto_pass)
