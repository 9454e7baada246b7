"""Generate tests/golden/*.npz from the REFERENCE ITSELF (run in the build container only).

    python -m oracle.make_golden

The reference (szbonaldo/FedMLP, mounted read-only at /root/reference) has no tests and no
golden vectors, so the fixtures are outputs of its own code on small seeded inputs:

  fedavg.npz      utils/FedAvg.py  FedAvg / FedAvg_proto / FedAvg_tao, called directly
  aggregators.npz utils/FedAvg.py RSCFed / model_dist / FedAvg_rela, utils/FedNoRo.py DaAgg / model_dist
  tagging.npz     utils/local_training.py CosineSimilarityFast + utils/utils.py
                  max_m_indices / min_n_indices (incl. ties), called directly
  maskfill.npz    utils/local_training.py DatasetSplit_pseudo.__getitem__, called directly
  flow.npz        the UNMODIFIED LocalUpdate.train_FedMLP driven on CPU for the last stage-1
                  round and two stage-2 rounds of one client, with every tensor crossing the
                  hot-path seams recorded (features, similarities, selections, label/mask fill,
                  per-step logits / loss / logit-gradients, prototypes, t)

CPU execution of the reference needs three environment shims, none of which changes its
arithmetic: Tensor.cuda()/Module.cuda() become no-ops, torch.cuda.FloatTensor is
torch.FloatTensor, and DataLoader is forced to num_workers=0 (the sandbox has no /dev/shm
budget for 8 workers; order and contents are unaffected).  TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import contextlib
import copy
import io
import os
import sys
import types
from pathlib import Path

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from oracle import ref_loader

GOLDEN_DIR = Path(__file__).resolve().parent.parent / "tests" / "golden"


def _np(t):
    return t.detach().cpu().numpy().copy() if isinstance(t, torch.Tensor) else np.array(t)


# ----------------------------------------------------------------------------------- fedavg
def make_fedavg(ref):
    g = torch.Generator().manual_seed(1037)
    shapes = OrderedShapes = [("features.conv0.weight", (7, 3, 5)), ("features.norm0.weight", (13,)),
                              ("features.norm0.running_mean", (13,)), ("features.norm0.num_batches_tracked", ()),
                              ("classifier.weight", (5, 33)), ("classifier.bias", (5,)),
                              ("features.norm1.num_batches_tracked", ())]
    K = 4
    base = {}
    for name, shp in shapes:
        base[name] = torch.randn(shp, generator=g) if "num_batches" not in name else None
    clients = []
    for k in range(K):
        sd = {}
        for name, shp in shapes:
            if "num_batches" in name:
                sd[name] = torch.tensor(100 + 7 * k, dtype=torch.int64)
            else:
                sd[name] = base[name] + 0.02 * torch.randn(shp, generator=g)
        clients.append(sd)
    from collections import OrderedDict
    clients = [OrderedDict(c) for c in clients]
    out = {}
    names = [n for n, _ in shapes]
    out["names"] = np.array(names)
    for k, sd in enumerate(clients):
        for n in names:
            out[f"in/{k}/{n}"] = _np(sd[n])
    for tag, weights in (("int", [5000, 4999, 5001, 1234]), ("float", [0.25, 0.125, 0.5, 0.125]),
                         ("float_odd", [0.3, 0.1, 0.45, 0.15])):
        res = ref.FedAvg.FedAvg(clients, weights)
        out[f"weights/{tag}"] = np.array(weights, dtype=np.float64)
        for n in names:
            out[f"out/{tag}/{n}"] = _np(res[n])
            out[f"outdtype/{tag}/{n}"] = np.array(str(res[n].dtype))
    # prototypes / tao
    C, D = 5, 24
    protos = [torch.randn(2 * C, D, generator=g) for _ in range(K)]
    weight = [5000, 4999, 5001, 1234]
    active_lists = [[0, 3], [1], [2], [], [0, 1, 2, 3]]       # class 3 has no annotating client -> NaN rows
    res = ref.FedAvg.FedAvg_proto(protos, weight, active_lists)
    out["proto/in"] = np.stack([_np(p) for p in protos])
    out["proto/weight"] = np.array(weight)
    out["proto/lists"] = np.array([",".join(map(str, l)) for l in active_lists])
    out["proto/out"] = _np(res)
    taos = [torch.rand(C, generator=g).double().numpy().copy() for _ in range(K)]
    neg_lists = [[1, 2], [0, 2, 3], [0, 1, 3], [0, 1, 2, 3], []]
    out["tao/in"] = np.stack(taos)
    out["tao/lists"] = np.array([",".join(map(str, l)) for l in neg_lists])
    out["tao/out_lists"] = ref.FedAvg.FedAvg_tao(taos, weight, neg_lists)
    out["tao/out_plain"] = ref.FedAvg.FedAvg_tao(taos, weight)
    np.savez_compressed(GOLDEN_DIR / "fedavg.npz", **out)


# ----------------------------------------------------------------------------------- other aggregators
def make_aggregators(ref):
    """utils/FedAvg.py RSCFed / model_dist / FedAvg_rela and utils/FedNoRo.py DaAgg / model_dist."""
    from collections import OrderedDict
    g = torch.Generator().manual_seed(77)
    shapes = [("conv.weight", (6, 3, 3, 3)), ("bn.weight", (6,)), ("bn.running_var", (6,)),
              ("bn.num_batches_tracked", ()), ("fc.weight", (5, 2500)), ("fc.bias", (5,))]
    K = 6
    base = {n: torch.randn(s, generator=g) for n, s in shapes if "num_batches" not in n}
    clients = []
    for k in range(K):
        sd = OrderedDict()
        for n, s in shapes:
            sd[n] = torch.tensor(50 + k, dtype=torch.int64) if "num_batches" in n else base[n] + 0.05 * torch.randn(s, generator=g)
        clients.append(sd)
    out = {"names": np.array([n for n, _ in shapes])}
    for k, sd in enumerate(clients):
        for n in sd:
            out[f"in/{k}/{n}"] = _np(sd[n])
    dict_len = [5000, 4000, 6000, 3000, 5500, 4500]
    out["dict_len"] = np.array(dict_len)
    # model_dist (FedNoRo version skips int tensors)
    out["model_dist/0_1"] = np.array(ref.FedNoRo.model_dist(clients[0], clients[1]))
    out["model_dist/2_5"] = np.array(ref.FedNoRo.model_dist(clients[2], clients[5]))
    # DaAgg
    clean, noisy = [0, 2, 3], [1, 4, 5]
    res = ref.FedNoRo.DaAgg(clients, dict_len, clean, noisy)
    out["daagg/clean"], out["daagg/noisy"] = np.array(clean), np.array(noisy)
    for n in res:
        out[f"daagg/out/{n}"] = _np(res[n])
        out[f"daagg/dtype/{n}"] = np.array(str(res[n].dtype))
    # RSCFed works on float-only dicts in the reference (its model_dist calls torch.norm on every key)
    fclients = [OrderedDict((n, v) for n, v in sd.items() if v.dtype == torch.float32) for sd in clients]
    DMA = [[0, 1, 2], [3, 4, 5], [1, 3, 5], [0, 2, 4]]
    res = ref.FedAvg.RSCFed(DMA, fclients, 3, dict_len, 4)
    out["rscfed/dma"] = np.array(DMA)
    out["rscfed/model_dist_0_1"] = np.array(ref.FedAvg.model_dist(fclients[0], fclients[1]))
    for n in res:
        out[f"rscfed/out/{n}"] = _np(res[n])
    # FedAvg_rela
    C, D = 5, 24
    protos = [torch.randn(C, D, generator=g) for _ in range(K)]
    lists = [[0, 3], [1], [2, 4, 5], [], [0, 1, 2, 3, 4, 5]]
    out["rela/in"] = np.stack([_np(p) for p in protos])
    out["rela/lists"] = np.array([",".join(map(str, l)) for l in lists])
    with np.errstate(all="ignore"):
        out["rela/out"] = _np(ref.FedAvg.FedAvg_rela(protos, dict_len, lists))
    np.savez_compressed(GOLDEN_DIR / "aggregators.npz", **out)


# ----------------------------------------------------------------------------------- tagging
def make_tagging(ref):
    g = torch.Generator().manual_seed(2024)
    N, D, C = 257, 40, 5
    feats = torch.relu(torch.randn(N, D, generator=g))
    feats[17] = feats[5]          # duplicate rows -> exactly equal similarities (tie handling)
    feats[200] = feats[5]
    feats[33] = feats[101]
    proto = torch.relu(torch.randn(2 * C, D, generator=g)) + 0.1
    out = {"feat": _np(feats), "proto": _np(proto)}
    model = ref.local_training.CosineSimilarityFast()
    for c in range(C):
        c0 = model(feats, torch.unsqueeze(proto[2 * c], dim=0))
        c1 = model(feats, torch.unsqueeze(proto[2 * c + 1], dim=0))
        out[f"cos0/{c}"] = _np(c0)
        out[f"cos1/{c}"] = _np(c1)
        sim = (c0 - c1)
        out[f"sim/{c}"] = _np(sim)
        lst = sim.tolist()
        for n in (0, 1, 3, 10, 40):
            out[f"max/{c}/{n}"] = np.array(ref.utils.max_m_indices(lst, n), dtype=np.int64)
            out[f"min/{c}/{n}"] = np.array(ref.utils.min_n_indices(lst, n), dtype=np.int64)
    # a hand-made list with many ties and signed zeros
    vals = [0.5, -0.25, 0.5, 0.0, -0.0, 0.125, -0.25, 0.5, -1.0, 0.125, -1.0, 0.0]
    out["ties/vals"] = np.array(vals, dtype=np.float32)
    for n in range(0, len(vals) + 1):
        out[f"ties/max/{n}"] = np.array(ref.utils.max_m_indices(vals, n), dtype=np.int64)
        out[f"ties/min/{n}"] = np.array(ref.utils.min_n_indices(vals, n), dtype=np.int64)
    np.savez_compressed(GOLDEN_DIR / "tagging.npz", **out)


def make_pool(ref):
    """Fused pooling + tagging: the feature the reference tags is the tail of the model
    forward, `features = self.features(x); out = F.relu(features); out = F.adaptive_avg_pool2d(out,
    (1, 1)); out = torch.flatten(out, 1)` (torchvision densenet.py DenseNet.forward, which the authors
    patched to return (out, logits); not vendored).  The tail is run here through the
    installed torchvision DenseNet class itself (features = Identity, classifier = Identity) and the
    similarities through the reference's CosineSimilarityFast."""
    import torchvision
    g = torch.Generator().manual_seed(77)
    B, D, H, W, C = 9, 40, 7, 7, 5
    net = torchvision.models.DenseNet(growth_rate=4, block_config=(1,), num_init_features=8, num_classes=C)
    net.features = nn.Identity()
    net.classifier = nn.Identity()
    net.eval()
    fmap = torch.randn(B, D, H, W, generator=g)
    fmap[3, 7] = -1.0                                  # a channel that pools to exactly 0
    with torch.no_grad():
        feat = net(fmap.clone())                       # relu (in place) + avg-pool + flatten
        feat_norelu = F.adaptive_avg_pool2d(fmap, (1, 1)).flatten(1)   # EfficientNet tail: no relu
    proto = torch.relu(torch.randn(2 * C, D, generator=g)) + 0.1
    out = {"fmap": _np(fmap), "feat": _np(feat), "feat_norelu": _np(feat_norelu), "proto": _np(proto)}
    model = ref.local_training.CosineSimilarityFast()
    for c in range(C):
        for tag, f in (("", feat), ("_norelu", feat_norelu)):
            c0 = model(f, torch.unsqueeze(proto[2 * c], dim=0))
            c1 = model(f, torch.unsqueeze(proto[2 * c + 1], dim=0))
            out[f"sim{tag}/{c}"] = _np(c0 - c1)
    np.savez_compressed(GOLDEN_DIR / "pool.npz", **out)


def make_eval(ref):
    """Evaluation metrics: the reference's own `globaltest` / `classtest`
    (utils/evaluations.py:15-73, :89-140; sklearn + utils/multilabel_metrixs.py) on a synthetic test set
    with a tiny CPU model.  Shim: DataLoader workers -> 0 (same batches, same order)."""
    import importlib
    ev = importlib.import_module("utils.evaluations")
    real_loader = ev.DataLoader
    ev.DataLoader = lambda *a, **k: real_loader(*a, **{**k, "num_workers": 0})

    class EvalSet(torch.utils.data.Dataset):
        def __init__(self, n, c, dim, seed):
            g = torch.Generator().manual_seed(seed)
            self.targets = (torch.rand(n, c, generator=g) < torch.linspace(0.05, 0.4, c)).float().numpy()
            self.targets[:3] = 0; self.targets[3:6] = 1          # every class has both labels
            shift = torch.randn(c, dim, generator=g)
            self.x = (0.7 * torch.randn(n, dim, generator=g) + torch.from_numpy(self.targets) @ shift).float()
            self.x[50:60] = self.x[40:50]                          # duplicated inputs -> tied probabilities
            self.x[200:203] = self.x[7]

        def __getitem__(self, i):
            return {"image": self.x[i], "target": self.targets[i]}

        def __len__(self):
            return len(self.targets)

    class Net(nn.Module):
        def __init__(self, dim, c):
            super().__init__()
            self.fc = nn.Linear(dim, c)

        def forward(self, x):
            return x, self.fc(x)

    torch.manual_seed(11)
    n, c, dim = 611, 5, 12
    ds, net = EvalSet(n, c, dim, 5), Net(dim, c)
    args = types.SimpleNamespace(batch_size=16, device="cpu", n_classes=c)
    with contextlib.redirect_stdout(io.StringIO()):
        res = ev.globaltest(net, ds, args)
        cls = [ev.classtest(net, ds, args, i) for i in range(c)]
    with torch.no_grad():
        logits = net(ds.x)[1]
    out = {"logits": _np(logits), "probs": _np(torch.sigmoid(logits)), "labels": ds.targets.astype(np.float32)}
    for k, v in res.items():
        out[f"globaltest/{k}"] = np.array(float(v), dtype=np.float64)
    for i, d in enumerate(cls):
        for k, v in d.items():
            out[f"classtest/{i}/{k}"] = np.array(float(v), dtype=np.float64)
    # per-class sklearn values on a hand-made score vector with heavy ties
    from sklearn.metrics import average_precision_score, roc_curve, auc
    ys = np.array([1, 0, 1, 1, 0, 0, 1, 0, 0, 1, 0, 1], dtype=np.float32)
    sc = np.array([0.9, 0.9, 0.5, 0.5, 0.5, 0.1, 0.1, 0.7, 0.7, 0.7, 0.3, 0.3], dtype=np.float32)
    fpr, tpr, _ = roc_curve(ys, sc, pos_label=1)
    out["ties/y"], out["ties/p"] = ys, sc
    out["ties/ap"] = np.array(average_precision_score(ys, sc)); out["ties/auc"] = np.array(auc(fpr, tpr))
    ev.DataLoader = real_loader
    np.savez_compressed(GOLDEN_DIR / "eval.npz", **out)


# ----------------------------------------------------------------------------------- synthetic dataset / model
class SynthDataset(torch.utils.data.Dataset):
    """Same sample contract as dataset/all_dataset.py:23-41 (two-view dict, numpy target row)."""

    def __init__(self, n, c, dim, seed):
        g = torch.Generator().manual_seed(seed)
        self.targets = (torch.rand(n, c, generator=g) < 0.3).float().numpy().astype(np.float32)
        for col in range(c):                      # every class has positives (LocalUpdate divides by the count)
            self.targets[col::c, col][:2] = 1.0
        shift = torch.randn(c, dim, generator=g)
        base = torch.randn(n, dim, generator=g)
        self.view1 = (base + torch.from_numpy(self.targets) @ shift).float()
        self.view2 = (self.view1 + 0.05 * torch.randn(n, dim, generator=g)).float()

    def __getitem__(self, index):
        return {"image_aug_1": self.view1[index], "image_aug_2": self.view2[index],
                "target": self.targets[index], "index": index, "image_id": str(index)}

    def __len__(self):
        return len(self.targets)


TRACE = []   # (kind, payload) in program order


class TinyNet(nn.Module):
    """net(x) -> (feature, logits), the contract the hot path relies on."""

    def __init__(self, dim, feat_dim, c):
        super().__init__()
        self.fc1 = nn.Linear(dim, feat_dim)
        self.fc2 = nn.Linear(feat_dim, c)
        self.role = "student"

    def forward(self, x):
        f = torch.relu(self.fc1(x))
        z = self.fc2(f)
        rec = {"role": self.role, "training": self.training, "grad": z.requires_grad,
               "feature": _np(f), "logits": _np(z), "dlogits": None}
        TRACE.append(("forward", rec))
        if z.requires_grad:
            z.register_hook(lambda g, rec=rec: rec.__setitem__("dlogits", _np(g)))
        return f, z

    def __deepcopy__(self, memo):
        # train_FedMLP deep-copies the incoming net as the frozen global model (:908, :1018)
        new = TinyNet(self.fc1.in_features, self.fc1.out_features, self.fc2.out_features)
        new.load_state_dict(copy.deepcopy(self.state_dict()))
        new.role = "global"
        new.train(self.training)
        return new


@contextlib.contextmanager
def cpu_shims(lt):
    saved = (torch.Tensor.cuda, nn.Module.cuda, getattr(torch.cuda, "FloatTensor", None), lt.DataLoader,
             torch.Tensor.backward)
    torch.Tensor.cuda = lambda self, *a, **k: self
    nn.Module.cuda = lambda self, *a, **k: self
    torch.cuda.FloatTensor = torch.FloatTensor
    real_loader = lt.DataLoader

    def loader(*a, **k):
        k["num_workers"] = 0
        return real_loader(*a, **k)

    lt.DataLoader = loader
    real_backward = torch.Tensor.backward

    def backward(self, *a, **k):
        TRACE.append(("loss", float(self.detach())))
        return real_backward(self, *a, **k)

    torch.Tensor.backward = backward
    try:
        yield
    finally:
        torch.Tensor.cuda, nn.Module.cuda, fl, lt.DataLoader, torch.Tensor.backward = saved
        if fl is not None:
            torch.cuda.FloatTensor = fl


@contextlib.contextmanager
def recorders(lt):
    """Record what crosses the seams of the hot path without touching the reference code."""
    cos_fwd = lt.CosineSimilarityFast.forward
    max_m, min_n = lt.max_m_indices, lt.min_n_indices
    la_fwd = lt.LogitAdjust_Multilabel.forward
    pseudo_get = lt.DatasetSplit_pseudo.__getitem__
    split_get = lt.DatasetSplit.__getitem__

    def cos(self, x1, x2):
        out = cos_fwd(self, x1, x2)
        TRACE.append(("cos", {"x1": _np(x1), "x2": _np(x2), "out": _np(out)}))
        return out

    def mx(lst, n):
        r = max_m(lst, n)
        TRACE.append(("max_m", {"values": np.array(lst, dtype=np.float32), "n": n, "out": np.array(r, dtype=np.int64)}))
        return r

    def mn(lst, n):
        r = min_n(lst, n)
        TRACE.append(("min_n", {"values": np.array(lst, dtype=np.float32), "n": n, "out": np.array(r, dtype=np.int64)}))
        return r

    def la(self, x, target):
        out = la_fwd(self, x, target)
        TRACE.append(("bce", {"p": _np(x), "target": _np(target), "out": _np(out)}))
        return out

    def pget(self, item):
        sample, idx, distill = pseudo_get(self, item)
        TRACE.append(("pseudo_item", {"idx": int(idx), "target": np.array(sample["target"]).copy(), "distill": _np(distill)}))
        return sample, idx, distill

    def sget(self, item):
        sample, idx, act = split_get(self, item)
        TRACE.append(("split_item", {"idx": int(idx), "target": np.array(sample["target"]).copy()}))
        return sample, idx, act

    lt.CosineSimilarityFast.forward = cos
    lt.max_m_indices, lt.min_n_indices = mx, mn
    lt.LogitAdjust_Multilabel.forward = la
    lt.DatasetSplit_pseudo.__getitem__ = pget
    lt.DatasetSplit.__getitem__ = sget
    try:
        yield
    finally:
        lt.CosineSimilarityFast.forward = cos_fwd
        lt.max_m_indices, lt.min_n_indices = max_m, min_n
        lt.LogitAdjust_Multilabel.forward = la_fwd
        lt.DatasetSplit_pseudo.__getitem__ = pseudo_get
        lt.DatasetSplit.__getitem__ = split_get


def make_flow_args(**over):
    a = types.SimpleNamespace(batch_size=16, annotation_num=1, n_classes=5, n_clients=5, local_ep=1,
                              base_lr=3e-3, device="cpu", U=0.7, L=0.3, rounds_FedMLP_stage1=3,
                              clean_threshold=0.12, noise_threshold=0.2, dataset="ICH", exp="FedMLP")
    a.__dict__.update(over)
    return a


def _dump_trace(prefix, out):
    """Flatten TRACE into npz keys `prefix/<i>/<kind>/<field>`; returns the kind sequence."""
    kinds = []
    for i, (kind, rec) in enumerate(TRACE):
        kinds.append(kind)
        if kind == "loss":
            out[f"{prefix}/{i}/loss"] = np.array(rec, dtype=np.float64)
            continue
        for k, v in rec.items():
            if v is None:
                continue
            out[f"{prefix}/{i}/{kind}/{k}"] = np.array(v)
    out[f"{prefix}/kinds"] = np.array(kinds)
    TRACE.clear()


def make_flow(ref):
    lt = ref.local_training
    N, C, DIM, FEAT = 112, 5, 12, 32
    CLIENT = 2                                 # client 2 annotates class 2 (main.py:74-77)
    torch.manual_seed(7)
    np.random.seed(7)
    import random
    random.seed(7)
    ds = SynthDataset(N, C, DIM, seed=99)
    targets_true = ds.targets.copy()
    idxs = list(range(4, N - 4))                # a strict subset, like a dict_users partition
    # main.py:58-66: all positives of every class are candidates for hiding (p_pos_1 = 0)
    rows, cols = np.where(ds.targets == 1)
    class_neg_idx = [rows[np.where(cols == i)[0]] for i in range(C)]
    args = make_flow_args()
    out = {"meta/N": np.array(N), "meta/C": np.array(C), "meta/FEAT": np.array(FEAT),
           "meta/client": np.array(CLIENT), "meta/idxs": np.array(idxs), "meta/targets_true": targets_true,
           "meta/batch_size": np.array(args.batch_size), "meta/L": np.array(args.L), "meta/U": np.array(args.U),
           "meta/clean_threshold": np.array(args.clean_threshold), "meta/noise_threshold": np.array(args.noise_threshold)}
    net = TinyNet(DIM, FEAT, C)
    sink = io.StringIO()
    with cpu_shims(lt), recorders(lt), contextlib.redirect_stdout(sink):
        local = lt.LocalUpdate(args, CLIENT, copy.deepcopy(ds), idxs, class_neg_idx, class_neg_idx,
                               active_class_list=[CLIENT], student=None, teacher_neg=None, teacher_act=None)
        TRACE.clear()
        tao = [0] * C
        # ---- last stage-1 round (rnd == stage1-1): losses + first prototypes / t (:907-1002)
        rnd = args.rounds_FedMLP_stage1 - 1
        work = copy.deepcopy(net)
        work.role = "student"
        # main.py:180-184: from round stage1-1 on the driver passes the client's class lists
        neg_in = [c for c in range(C) if c != CLIENT]
        ret = local.train_FedMLP(rnd, tao, [], None, negetive_class_list=neg_in, active_class_list_client_i=[CLIENT], net=work)
        w_local, loss_mean, _, _, neg_list, act_list, t_local, proto_local = ret
        out["s1/loss_mean"] = np.array(loss_mean)
        out["s1/neg_list"] = np.array(neg_list)
        out["s1/act_list"] = np.array(act_list)
        out["s1/t"] = np.array(t_local)
        out["s1/proto"] = _np(proto_local)
        _dump_trace("s1", out)
        # the server would aggregate prototypes of all clients; for one client use its own rows for
        # the active class and seeded non-degenerate rows for the others (only rows 2c, 2c+1 of the
        # *missing* classes are read by the tagger)
        g = torch.Generator().manual_seed(5)
        proto_glob = torch.relu(torch.randn(2 * C, FEAT, generator=g)) + 0.05
        proto_glob[2 * CLIENT] = proto_local[2 * CLIENT]
        proto_glob[2 * CLIENT + 1] = proto_local[2 * CLIENT + 1]
        out["proto_glob"] = _np(proto_glob)
        # ---- two stage-2 rounds (:1006-1256)
        for r in range(2):
            rnd = args.rounds_FedMLP_stage1 + r
            work2 = copy.deepcopy(work)
            work2.role = "student"
            ret = local.train_FedMLP(rnd, tao, proto_glob, None, negetive_class_list=list(neg_list),
                                     active_class_list_client_i=list(act_list), net=work2)
            w_local, loss_mean, _, _, neg2, act2, t2, proto2 = ret
            out[f"s2_{r}/loss_mean"] = np.array(loss_mean)
            out[f"s2_{r}/t"] = np.array(t2)
            out[f"s2_{r}/proto"] = _np(proto2)
            for j, lst in enumerate(local.traindata_idx):
                out[f"s2_{r}/traindata_idx/{j}"] = np.array(lst, dtype=np.float64)
            for j, lst in enumerate(local.idxss):
                out[f"s2_{r}/idxss/{j}"] = np.array(sorted(lst), dtype=np.int64)
            _dump_trace(f"s2_{r}", out)
            work = work2
    np.savez_compressed(GOLDEN_DIR / "flow.npz", **out)


# ----------------------------------------------------------------------------------- mask fill
def make_maskfill(ref):
    lt = ref.local_training
    N, C = 40, 5
    ds = SynthDataset(N, C, 4, seed=3)
    idxs = list(range(3, 35))
    args = types.SimpleNamespace(annotation_num=1, n_classes=C)
    active, negative = [1], [0, 2, 3, 4]
    traindata_idx = [[5.0, 9.0], [7.0], [], [10.0, 11.0], [12.0, 5.0], [], [30.0], [31.0, 7.0]]
    pseudo = lt.DatasetSplit_pseudo(copy.deepcopy(ds), idxs, 1, args, active, negative, traindata_idx)
    tgt, dis, ids = [], [], []
    for item in range(len(pseudo)):
        sample, idx, distill = pseudo[item]
        tgt.append(np.array(sample["target"]).copy())
        dis.append(_np(distill))
        ids.append(idx)
    np.savez_compressed(GOLDEN_DIR / "maskfill.npz", targets_true=ds.targets, idxs=np.array(idxs),
                        active=np.array(active), negative=np.array(negative),
                        traindata_idx=np.array([",".join(str(v) for v in l) for l in traindata_idx]),
                        out_target=np.stack(tgt), out_distill=np.stack(dis), out_idx=np.array(ids))


def main():
    if not ref_loader.available():
        sys.exit("reference not mounted; golden fixtures can only be regenerated in the build container")
    GOLDEN_DIR.mkdir(parents=True, exist_ok=True)
    ref = ref_loader.load()
    torch.set_num_threads(1)   # fixed reduction order for the recorded sums
    make_fedavg(ref)
    make_aggregators(ref)
    make_tagging(ref)
    make_pool(ref)
    make_eval(ref)
    make_maskfill(ref)
    make_flow(ref)
    for p in sorted(GOLDEN_DIR.glob("*.npz")):
        print(p.name, os.path.getsize(p), "bytes")


if __name__ == "__main__":
    main()
