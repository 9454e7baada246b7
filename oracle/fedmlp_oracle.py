"""CPU restatement of the FedMLP per-round hot path — TEST INFRASTRUCTURE ONLY.

This module is the *oracle* for fedmlp_b200's CUDA kernels.  It restates, with stock PyTorch /
numpy CPU ops, what szbonaldo/FedMLP computes on the path tag -> prototypes -> loss -> FedAvg.
Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline / `--impl reference` legs may
import it; nothing under fedmlp_b200/ does (the product path has no CPU fallback).

Pinning: the reference ships no tests or golden vectors.  The oracle is
pinned instead against outputs of the reference's own code run in the build container:
oracle/make_golden.py imports /root/reference (FedAvg*, CosineSimilarityFast, max_m_indices,
min_n_indices, DatasetSplit_pseudo, LogitAdjust_Multilabel) and drives the unmodified
LocalUpdate.train_FedMLP on CPU, and writes tests/golden/*.npz; tests/test_oracle_golden.py
checks every function below against those files.

Each function cites the reference lines it follows (paths relative to /root/reference).
The arithmetic that is not in the reference repo is stock PyTorch (reference pins
torch==1.12.1+cu116, requirements.txt:100; this image has torch 2.11): sigmoid,
F.binary_cross_entropy (log clamped at -100, backward eps 1e-12), mm, norm, sum.
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F


# ------------------------------------------------------------------------------------ a1-a3
def fedavg(w, dict_len):
    """utils/FedAvg.py:7-14 (== Fed_w :16-23).  Per key: scale client 0, add the other scaled
    clients one after the other, divide by the total weight.  int64 entries stay int64 until
    the true division (int weights) and come out float32."""
    total = sum(dict_len)
    out = OrderedDict()
    for key, first in w[0].items():
        acc = first * dict_len[0]
        for i in range(1, len(w)):
            acc = acc + w[i][key] * dict_len[i]   # `+=` on a fresh tensor: same values
        out[key] = acc / total
    return out


def fedavg_proto(prototypes, weight, class_active_client_list):
    """utils/FedAvg.py:72-93.  Per class, weighted mean of rows 2c / 2c+1 over the clients that
    annotate the class (list order), `P*w + acc` starting from zeros; an empty client list
    divides 0 by 0."""
    n_rows, dim = len(prototypes[0]), len(prototypes[0][0])
    avg = torch.zeros((n_rows, dim))
    w_arr = np.array(weight)
    for cls, clients in enumerate(class_active_client_list):
        acc0 = torch.zeros_like(prototypes[0][0])
        acc1 = torch.zeros_like(prototypes[0][0])
        for cid in clients:
            acc0 = prototypes[cid][2 * cls] * weight[cid] + acc0
            acc1 = prototypes[cid][2 * cls + 1] * weight[cid] + acc1
        denom = np.sum(w_arr[clients])
        avg[2 * cls] = acc0 / denom
        avg[2 * cls + 1] = acc1 / denom
    return avg


def fedavg_tao(t, weight, class_client_list=None):
    """utils/FedAvg.py:51-70 (float64 numpy).  main.py:223 passes the per-class lists of clients
    for which the class is *missing*; a class nobody misses gets 1.0."""
    n = len(t[0])
    avg = np.zeros(n, dtype=np.float64)
    if class_client_list is None:
        for i, tao in enumerate(t):
            avg += tao * float(weight[i])
        return avg / float(sum(weight))
    for cls, clients in enumerate(class_client_list):
        wsum = 0.0
        for i, tao in enumerate(t):
            if i in clients:
                avg[cls] += tao[cls] * float(weight[i])
                wsum += float(weight[i])
        avg[cls] = 1.0 if len(clients) == 0 else avg[cls] / wsum
    return avg


# ------------------------------------------------------------------------------------ §8f.3 other aggregators
def fedavg_rela(prototypes, weight, class_active_client_list):
    """utils/FedAvg.py:95-103: FedAvg_proto with one row per class."""
    avg = torch.zeros((len(prototypes[0]), len(prototypes[0][0])))
    w_arr = np.array(weight)
    for cls, clients in enumerate(class_active_client_list):
        acc = torch.zeros_like(prototypes[0][0])
        for cid in clients:
            acc = prototypes[cid][cls] * weight[cid] + acc
        avg[cls] = acc / np.sum(w_arr[clients])
    return avg


def model_dist(w_1, w_2):
    """utils/FedNoRo.py:106-115: per-key L2 norms of the difference (int tensors skipped), added in
    key order in fp32; returned as a Python float."""
    assert w_1.keys() == w_2.keys()
    total = torch.zeros(1).float()
    for key in w_1.keys():
        if "int" in str(w_1[key].dtype):
            continue
        total += torch.norm(w_1[key] - w_2[key])
    return total.item()


def rscfed(dma, w_locals, k, dict_len, m):
    """utils/FedAvg.py:25-40."""
    import math
    w_sub = []
    for group in dma:
        sel = [w_locals[i] for i in group]
        n_total = sum(dict_len[i] for i in group)
        w_avg = fedavg(sel, [1] * k)
        ws = [(dict_len[i] / n_total) * math.exp(-0.01 * (model_dist(w_locals[i], w_avg) / dict_len[i])) for i in group]
        w_sub.append(fedavg(sel, ws))
    return fedavg(w_sub, [1] * m)


def daagg(w, dict_len, clean_clients, noisy_clients):
    """utils/FedNoRo.py:84-103."""
    cw = np.array(dict_len)
    cw = cw / cw.sum()
    distance = np.zeros(len(dict_len))
    for n_idx in noisy_clients:
        distance[n_idx] = min(model_dist(w[n_idx], w[c_idx]) for c_idx in clean_clients)
    distance = distance / distance.max()
    cw = cw * np.exp(-distance)
    cw = cw / cw.sum()
    out = OrderedDict()
    for key, first in w[0].items():
        acc = first * cw[0]
        for i in range(1, len(w)):
            acc = acc + w[i][key] * cw[i]
        out[key] = acc
    return out


# ------------------------------------------------------------------------------------ a6
def cosine_similarity_fast(x1, x2):
    """utils/local_training.py:1417-1435 CosineSimilarityFast.forward: x1 [N,D], x2 [1,D] ->
    (x1 @ x2^T) * (1 / (|x1|_row (x) |x2|)), no epsilon, squeezed to [N]."""
    x2t = x2.t()
    dots = x1.mm(x2t)
    n1 = x1.norm(dim=1).unsqueeze(0).t()
    n2 = x2t.norm(dim=0).unsqueeze(0)
    return torch.squeeze(dots.mul(1 / n1.mm(n2)), dim=1)


def tag_similarity(features, prototype, missing_classes):
    """utils/local_training.py:1052-1058: per missing class cos(f, P[2c]) - cos(f, P[2c+1]).
    Returns {class: float32 tensor [N]}."""
    out = {}
    for cls in missing_classes:
        p0 = torch.unsqueeze(prototype[2 * cls], dim=0)
        p1 = torch.unsqueeze(prototype[2 * cls + 1], dim=0)
        out[cls] = cosine_similarity_fast(features, p0) - cosine_similarity_fast(features, p1)
    return out


def pooled_features(fmap, relu=True):
    """Tail of the reference's model forward, the `features` of `features, _ = net(images1)`
    (utils/local_training.py:1033-1036): torchvision densenet.py DenseNet.forward as patched by the
    authors (not vendored): relu -> adaptive_avg_pool2d(1) -> flatten.  EfficientNet's
    tail has no relu.  fmap [B, D, H, W] -> [B, D]."""
    x = torch.clamp_min(fmap, 0) if relu else fmap
    b, d = x.shape[0], x.shape[1]
    x = x.reshape(b, d, -1)
    return x.sum(dim=2) / x.shape[2]


def pool_tag(fmap, prototype, missing_classes, relu=True):
    """pooled features of one batch and their tagging similarities
    (utils/local_training.py:1033-1036 followed by :1052-1058 on those rows)."""
    feat = pooled_features(fmap, relu)
    return feat, tag_similarity(feat, prototype, missing_classes)


# ------------------------------------------------------------------------------------ a7
def top_positions_python(values, n, largest):
    """utils/utils.py:24-35 max_m_indices / min_n_indices, literally: Python's stable sort of
    (position, value) pairs; equal values keep increasing position in both directions."""
    pairs = sorted(enumerate(values), key=lambda pv: pv[1], reverse=largest)
    return [pos for pos, _ in pairs[:n]]


def top_positions(values, n, largest):
    """Same result as top_positions_python for NaN-free input, with a stable numpy argsort
    (used for the 55k / 85k-row cases where the interpreter sort takes seconds)."""
    v = np.asarray(values, dtype=np.float32)
    order = np.argsort(-v if largest else v, kind="stable")
    return order[:n].tolist()


def split_and_select(sim, clean_frac, noise_frac, valid=None, python_sort=False):
    """utils/local_training.py:1061-1072 for one class: sign split at 0, counts
    int(frac * len(side)), then the m largest / k smallest similarities.
    `valid` restricts the candidates (later stage-2 rounds only look at untagged rows);
    positions returned are positions in the full `sim` vector.
    Returns dict(n_clean, n_noise, m, k, clean=[positions], noise=[positions])."""
    sim = np.asarray(sim, dtype=np.float32)
    cand = np.arange(len(sim)) if valid is None else np.nonzero(np.asarray(valid))[0]
    s = sim[cand]
    n_clean = int(np.sum(s >= 0))
    n_noise = int(np.sum(s < 0))
    m = int(1 * clean_frac * n_clean)
    k = int(1 * noise_frac * n_noise)
    pick = top_positions_python if python_sort else top_positions
    clean = pick(s.tolist() if python_sort else s, m, True)
    noise = pick(s.tolist() if python_sort else s, k, False)
    return dict(n_clean=n_clean, n_noise=n_noise, m=m, k=k,
                clean=[int(cand[p]) for p in clean], noise=[int(cand[p]) for p in noise])


# ------------------------------------------------------------------------------------ a8
def mask_fill(labels, dataset_idx, active_classes, missing_classes, traindata_idx):
    """DatasetSplit_pseudo.__getitem__ utils/local_training.py:1456-1477 for all rows:
    labels [N,C] original targets, dataset_idx [N]; traindata_idx = [clean_c0, noise_c0, ...]
    in `missing_classes` order.  Returns (target [N,C], distill_cls [N,C]) float32."""
    labels = np.array(labels, dtype=np.float32, copy=True)
    n, c = labels.shape
    target = labels.copy()
    distill = np.zeros((n, c), dtype=np.float32)
    for col in range(c):
        if col not in active_classes:
            target[:, col] = 0
    for i, cls in enumerate(missing_classes):
        clean = set(traindata_idx[2 * i])
        noise = set(traindata_idx[2 * i + 1])
        for r in range(n):
            d = dataset_idx[r]
            if d in clean or d in noise:
                if d in noise:
                    target[r, cls] = 1
            else:
                distill[r, cls] = 1
    return target, distill


def remaining_indices(all_idx, traindata_idx):
    """utils/local_training.py:1197-1204: per missing class, the local indices that are in
    neither list (returned sorted; the reference keeps them in set-iteration order)."""
    out = []
    for i in range(len(traindata_idx) // 2):
        tagged = set(traindata_idx[2 * i]) | set(traindata_idx[2 * i + 1])
        out.append(sorted(set(all_idx) - tagged))
    return out


# ------------------------------------------------------------------------------------ a4
def prototype_build(features, labels, logits, active_classes, t_classes, L, U, guard_empty,
                    batch=128, logits_are_probs=False):
    """utils/local_training.py:973-1000 (guard_empty=False) and :1208-1249 (guard_empty=True):
    walk the rows in batches of 4*batch_size=128, per active class add the per-batch sums of the
    label-0 / label-1 rows, count them, count p<L or p>U for the classes in `t_classes`,
    divide at the end.  Returns (proto [2C,D] float32, num_proto list[2C], t float64 [C])."""
    n, d = features.shape
    c = labels.shape[1]
    proto = torch.zeros((2 * c, d))
    num = [0] * (2 * c)
    t = np.array([0] * c)
    for b0 in range(0, n, batch):
        feat = features[b0:b0 + batch]
        lab = labels[b0:b0 + batch]
        probs = logits[b0:b0 + batch] if logits_are_probs else torch.sigmoid(logits[b0:b0 + batch])
        for cls in active_classes:
            idx0 = torch.where(lab[:, cls] == 0)[0]
            idx1 = torch.where(lab[:, cls] == 1)[0]
            num[2 * cls] += len(idx0)
            num[2 * cls + 1] += len(idx1)
            proto[2 * cls] = feat[idx0, :].sum(0) + proto[2 * cls]
            proto[2 * cls + 1] = feat[idx1, :].sum(0) + proto[2 * cls + 1]
        for cls in t_classes:
            t[cls] += torch.sum(torch.logical_or(probs[:, cls] < L, probs[:, cls] > U)).item()
    for cls in active_classes:
        for r in (2 * cls, 2 * cls + 1):
            if num[r] == 0 and guard_empty:
                continue
            proto[r] = proto[r] / num[r]   # 0/0 -> NaN when unguarded, like the reference
    return proto, num, t / n


# ------------------------------------------------------------------------------------ a9, a10
def bce_on_probs(p, y):
    """utils/FedNoRo.py:16-22 LogitAdjust_Multilabel.forward with its logit-adjust lines
    commented out == F.binary_cross_entropy(p, y, reduction='none')."""
    return F.binary_cross_entropy(p.clone(), y, weight=None, reduction="none")


def stage1_loss(z1, z2, z3, z4, y, active_classes, missing_classes, batch_size, annotation_num=None):
    """utils/local_training.py:933-963.  z1/z2 student logits of the two views (need grad),
    z3/z4 frozen global-model logits.  Returns the scalar loss tensor (autograd-capable)."""
    a = len(active_classes) if annotation_num is None else annotation_num
    p1, p2 = torch.sigmoid(z1), torch.sigmoid(z2)
    with torch.no_grad():
        p3, p4 = torch.sigmoid(z3), torch.sigmoid(z4)
    dis = ((p1 - p3) ** 2 + (p2 - p4) ** 2) / 2.0
    sup = (bce_on_probs(p1, y) + bce_on_probs(p2, y)) / 2.0
    loss_sup = sup[:, list(active_classes)].sum() / (batch_size * a)
    loss_dis = dis[:, list(missing_classes)].sum() / (batch_size * len(missing_classes))
    loss_unsup = F.mse_loss(p1[:, list(missing_classes)], p2[:, list(missing_classes)])
    return loss_sup + 0.0 * loss_unsup + loss_dis


def stage2_loss(z, zg, y, distill_cls, variant="sup"):
    """utils/local_training.py:1171-1188.  variant 'sup' is the live line :1188,
    'sup_dis' the commented alternative :1187."""
    sup_cls = (~distill_cls.bool()).float()
    p = torch.sigmoid(z)
    loss_sup = bce_on_probs(p, y)
    if variant == "sup":
        return (loss_sup * sup_cls).sum() / sup_cls.sum()
    with torch.no_grad():
        pg = torch.sigmoid(zg)
    loss_dis = (p - pg) ** 2
    return ((loss_sup * sup_cls).sum() + (loss_dis * distill_cls).sum()) / (sup_cls.sum() + distill_cls.sum())


def loss_and_grads(fn, *zs_and_rest, n_grad):
    """Helper: run `fn` with the first n_grad tensors requiring grad; returns (loss, grads...)."""
    zs = [z.detach().clone().requires_grad_(True) for z in zs_and_rest[:n_grad]]
    loss = fn(*zs, *zs_and_rest[n_grad:])
    grads = torch.autograd.grad(loss, zs)
    return (loss.detach(),) + tuple(g.detach() for g in grads)


# ------------------------------------------------------------------------------------ a5 + a7 flow
# ------------------------------------------------------------------------------------ f4 evaluation
def _binary_clf_curve(y_true, y_score):
    """sklearn.metrics._ranking._binary_clf_curve (scikit-learn 1.x, the dependency behind
    utils/evaluations.py:9-10; not vendored): stable sort by decreasing score, one point per distinct
    score, cumulative true / false positives."""
    y_true = (np.asarray(y_true) == 1)
    y_score = np.asarray(y_score)
    order = np.argsort(y_score, kind="mergesort")[::-1]
    y_score, y_true = y_score[order], y_true[order]
    distinct = np.where(np.diff(y_score))[0]
    idx = np.r_[distinct, y_true.size - 1]
    tps = np.cumsum(y_true, dtype=np.float64)[idx]
    fps = 1 + idx - tps
    return fps, tps


def average_precision(y_true, y_score):
    """sklearn average_precision_score (binary): -sum(diff(recall) * precision[:-1]) over the
    precision-recall curve (thresholds reversed, final point (recall 0, precision 1) appended)."""
    fps, tps = _binary_clf_curve(y_true, y_score)
    ps = tps + fps
    precision = np.zeros_like(tps)
    np.divide(tps, ps, out=precision, where=(ps != 0))
    recall = tps / tps[-1]
    precision = np.hstack((precision[::-1], 1.0))
    recall = np.hstack((recall[::-1], 0.0))
    return float(-np.sum(np.diff(recall) * precision[:-1]))


def roc_auc(y_true, y_score):
    """sklearn roc_curve + auc (utils/evaluations.py:60-63): trapezoid over (fpr, tpr) with the origin."""
    fps, tps = _binary_clf_curve(y_true, y_score)
    fps, tps = np.r_[0.0, fps], np.r_[0.0, tps]
    fpr, tpr = fps / fps[-1], tps / tps[-1]
    return float(np.sum(np.diff(fpr) * (tpr[1:] + tpr[:-1]) / 2.0))


def eval_metrics(all_probs, all_labels, threshold=0.5):
    """utils/evaluations.py:35-73 (everything after the inference loop) with utils/multilabel_metrixs.py
    restated on arrays.  all_probs float32 [N, C], all_labels 0/1 [N, C]."""
    all_probs = np.asarray(all_probs)
    y_true = np.asarray(all_labels)
    y_pred = all_probs > threshold
    n, c = y_true.shape
    aps = [average_precision(y_true[:, i], all_probs[:, i]) for i in range(c)]
    m_ap = torch.tensor(aps).mean()
    bacc = r = f1 = p = 0
    with np.errstate(divide="ignore", invalid="ignore"):
        for i in range(c):
            yt, yp = y_true.T[i], y_pred.T[i]
            tp = np.sum(np.logical_and(yt, yp))
            recall1 = tp / np.sum(yt)
            recall0 = np.sum(~np.logical_or(yt, yp)) / (yt.size - np.count_nonzero(yt))
            bacc += (recall0 + recall1) / 2
            r += tp / np.sum(yt)
            f1 += (2 * tp) / (np.sum(yt) + np.sum(yp))
            if np.sum(yp) != 0:
                p += tp / np.sum(yp)
    hamming = 0
    for i in range(n):
        hamming += np.size(y_true[i] == y_pred[i]) - np.count_nonzero(y_true[i] == y_pred[i])
    auroc = sum(roc_auc(y_true.T[i], all_probs.T[i]) for i in range(c)) / c
    return {"mAP": m_ap, "BACC": bacc / c, "R": r / c, "F1": f1 / c, "auc": auroc, "P": p / c,
            "hamming_loss": hamming / (n * c), "APs": aps}


class TaggingState:
    """Cross-round tagging state of one client (self.traindata_idx / self.idxss,
    utils/local_training.py:1025,1088-1089,1111-1112,1197-1204)."""

    def __init__(self, dataset_idx, missing_classes):
        self.dataset_idx = [int(i) for i in dataset_idx]
        self.missing = list(missing_classes)
        self.traindata_idx = [[] for _ in range(2 * len(self.missing))]

    def step(self, features, prototype, clean_frac, noise_frac, python_sort=False):
        """One stage-2 tagging pass over rows given in `dataset_idx` order.  Candidates of a
        class are the rows not yet in either of its lists."""
        sims = tag_similarity(features, prototype, self.missing)
        stats = []
        for i, cls in enumerate(self.missing):
            tagged = set(self.traindata_idx[2 * i]) | set(self.traindata_idx[2 * i + 1])
            valid = np.array([d not in tagged for d in self.dataset_idx])
            r = split_and_select(sims[cls].numpy(), clean_frac, noise_frac, valid=valid, python_sort=python_sort)
            self.traindata_idx[2 * i].extend(self.dataset_idx[p] for p in r["clean"])
            self.traindata_idx[2 * i + 1].extend(self.dataset_idx[p] for p in r["noise"])
            stats.append(r)
        return sims, stats


# ------------------------------------------------------------------------------------ synthetic inputs
ICH_PREVALENCE = [0.015, 0.176, 0.128, 0.174, 0.230]   # preprocess/ICH_process.py:45-46


def class_prevalence(c):
    return ICH_PREVALENCE if c == 5 else np.linspace(0.02, 0.20, c).tolist()


def synth_client(n, d, c, seed, signed=False):
    """Seeded synthetic client: features (post-ReLU for DenseNet, signed for EfficientNet),
    Bernoulli labels with dataset-like prevalences, logits N(0,2)."""
    g = torch.Generator().manual_seed(seed)
    base = torch.randn(n, d, generator=g)
    feats = base * torch.sigmoid(torch.randn(n, d, generator=g)) if signed else torch.relu(base)
    prev = torch.tensor(class_prevalence(c))
    labels = (torch.rand(n, c, generator=g) < prev).float()
    # make the features weakly label-dependent so prototypes separate, as in practice
    shift = torch.randn(c, d, generator=g) * 0.25
    feats = feats + torch.relu(labels @ shift) if not signed else feats + labels @ shift
    logits = torch.randn(n, c, generator=g) * 2.0
    return feats.contiguous(), labels.contiguous(), logits.contiguous()


def synth_prototypes(feats, labels):
    """Prototypes as masked means of the features for every class (what the server would
    aggregate from the clients that annotate each class)."""
    c = labels.shape[1]
    rows = []
    for cls in range(c):
        for v in (0.0, 1.0):
            sel = labels[:, cls] == v
            rows.append(feats[sel].mean(0) if sel.any() else torch.zeros(feats.shape[1]))
    return torch.stack(rows).contiguous()
