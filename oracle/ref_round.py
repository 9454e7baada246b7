"""One FedMLP round hot path on the host CPU, composed from the REFERENCE'S OWN callables.

TEST / MEASUREMENT INFRASTRUCTURE ONLY (bench.py `--impl reference` and `cpu_baseline`).

Where the reference's code is a callable it is called unmodified (imported through oracle/ref_loader.py
from /root/reference or from the vendored copies under oracle/_ref):
    utils/FedAvg.py        FedAvg :7-14, FedAvg_proto :72-93, FedAvg_tao :51-70
    utils/local_training.py CosineSimilarityFast :1417-1435, DatasetSplit_pseudo.__getitem__ :1456-1477
    utils/utils.py         max_m_indices / min_n_indices :24-35
    utils/FedNoRo.py       LogitAdjust_Multilabel.forward :16-22
Where it is an inline block of LocalUpdate.train_FedMLP (which needs a CUDA device and a CNN) the block is
restated literally around those calls, on the same tensors the CUDA path consumes:
    :1052-1065 similarities + sign split, :1066-1112 counts / top-fraction / traindata_idx,
    :1171-1188 stage-2 loss + backward, :1223-1249 prototypes + t.
The CNN forward (cuDNN, out of scope) is replaced by precomputed features / logits, exactly as in the GPU arm.
"""
from __future__ import annotations

import copy
import time
import types

import numpy as np
import torch

from . import ref_loader


class _MemDataset:
    """In-memory stand-in for dataset/all_dataset.py:23-41: samples are dicts with a `target` row."""

    def __init__(self, targets: np.ndarray):
        self.targets = targets

    def __getitem__(self, i):
        return {"target": self.targets[i].copy()}

    def __len__(self):
        return len(self.targets)


def load_reference():
    """The reference namespace + the one shim needed on a CPU: LogitAdjust_Multilabel.__init__ builds a
    torch.cuda.FloatTensor (utils/FedNoRo.py:12) that its forward never uses (:18-21 are commented out)."""
    ref = ref_loader.load()
    return ref


def make_criterion(ref, cls_num_list, num):
    crit = ref.FedNoRo.LogitAdjust_Multilabel.__new__(ref.FedNoRo.LogitAdjust_Multilabel)
    torch.nn.Module.__init__(crit)
    crit.weight = None
    return crit


def client_step(ref, cl, proto, args, clean_frac, noise_frac, L, U, batch_rows=128):
    """tag -> mask fill -> stage-2 loss fwd+bwd -> prototypes + t for one client (reference order)."""
    lt = ref.local_training
    feat, labels, idxs = cl["feat"], cl["labels"], cl["idxs"]
    negetive_class_list, active = cl["missing"], cl["active"]
    # ---- utils/local_training.py:1052-1058
    similarity = []
    model = lt.CosineSimilarityFast()
    for cls in negetive_class_list:
        proto_0, proto_1 = proto[2 * cls], proto[2 * cls + 1]
        sim = (model(feat, torch.unsqueeze(proto_0, dim=0)) - model(feat, torch.unsqueeze(proto_1, dim=0))).tolist()
        similarity.append(sim)
    # ---- :1061-1065
    clean_idx, noise_idx = [], []
    for i in range(len(negetive_class_list)):
        clean_idx.append(np.where(np.array(similarity[i]) >= 0)[0].tolist())
        noise_idx.append(np.where(np.array(similarity[i]) < 0)[0].tolist())
    # ---- :1066-1089 (first stage-2 round: append)
    class_idx = torch.tensor(idxs, dtype=torch.float32)
    traindata_idx = []
    for i, cls in enumerate(negetive_class_list):
        num_clean_cls = int(1 * clean_frac * len(clean_idx[i]))
        num_noise_cls = int(1 * noise_frac * len(noise_idx[i]))
        max_m_indices_list = np.array(lt.max_m_indices(similarity[i], num_clean_cls))
        min_n_indices_list = np.array(lt.min_n_indices(similarity[i], num_noise_cls))
        if len(max_m_indices_list) == 0 and len(max_m_indices_list) == 0:
            negcls_clean_train_idx = []
            negcls_noise_train_idx = []
        elif len(min_n_indices_list) == 0 and len(max_m_indices_list) != 0:
            negcls_noise_train_idx = []
            negcls_clean_train_idx = np.array(class_idx)[max_m_indices_list].tolist()
        elif len(min_n_indices_list) != 0 and len(max_m_indices_list) == 0:
            negcls_noise_train_idx = np.array(class_idx)[min_n_indices_list].tolist()
            negcls_clean_train_idx = []
        else:
            negcls_clean_train_idx = np.array(class_idx)[max_m_indices_list].tolist()
            negcls_noise_train_idx = np.array(class_idx)[min_n_indices_list].tolist()
        traindata_idx.append(negcls_clean_train_idx)
        traindata_idx.append(negcls_noise_train_idx)
    # ---- DatasetSplit_pseudo.__getitem__ :1456-1477 for every local sample (what the DataLoader workers do)
    ds = lt.DatasetSplit_pseudo(_MemDataset(cl["targets_np"]), idxs, cl["client_id"], args, active, negetive_class_list,
                                traindata_idx)
    tgt = np.empty((len(idxs), args.n_classes), dtype=np.float32)
    dis = torch.empty((len(idxs), args.n_classes))
    for item in range(len(ds)):
        sample, _, distill_cls = ds[item]
        tgt[item] = sample["target"]
        dis[item] = distill_cls
    # ---- :1171-1188 over the client's rows (one big batch: the arithmetic per row is the same)
    z = cl["logits"].clone().requires_grad_(True)
    distill = dis
    sup_cls = (~distill.bool()).float()
    criterion = make_criterion(ref, None, len(idxs))
    logits1 = torch.sigmoid(z)
    with torch.no_grad():
        logits2 = torch.sigmoid(cl["zg"])
    loss_sup = criterion(logits1, torch.from_numpy(tgt))
    loss_dis = torch.nn.MSELoss(reduction="none")(logits1, logits2)   # computed and unused, as at :1185
    loss = (loss_sup * sup_cls).sum() / sup_cls.sum()
    loss.backward()
    # ---- :1223-1249 prototypes + t, batches of 4 * batch_size
    C = args.n_classes
    proto_out = torch.zeros((C * 2, feat.shape[1]))
    num_proto = [0] * C * 2
    t = np.array([0] * C)
    with torch.no_grad():
        for r0 in range(0, len(idxs), batch_rows):
            feature = cl["feat2"][r0:r0 + batch_rows]
            lab = labels[r0:r0 + batch_rows]
            probs = torch.sigmoid(cl["logits2"][r0:r0 + batch_rows])
            for cls in active:
                idx0 = torch.where(lab[:, cls] == 0)[0]
                idx1 = torch.where(lab[:, cls] == 1)[0]
                num_proto[2 * cls] += len(idx0)
                num_proto[2 * cls + 1] += len(idx1)
                proto_out[2 * cls] += feature[idx0, :].sum(0)
                proto_out[2 * cls + 1] += feature[idx1, :].sum(0)
            for cls in negetive_class_list:
                t[cls] += torch.sum(torch.logical_or(probs[:, cls] < L, probs[:, cls] > U)).item()
    for cls in active:
        if num_proto[2 * cls] != 0:
            proto_out[2 * cls] = proto_out[2 * cls] / num_proto[2 * cls]
        if num_proto[2 * cls + 1] != 0:
            proto_out[2 * cls + 1] = proto_out[2 * cls + 1] / num_proto[2 * cls + 1]
    t = t / len(idxs)
    return dict(traindata_idx=traindata_idx, loss=float(loss.detach()), dz=z.grad, proto=proto_out, t=t,
                similarity=similarity, target=tgt, distill=dis)


def make_clients(n_clients, rows, D, C, seed=1037, first_client=0, signed=False):
    from . import fedmlp_oracle as O

    clients = []
    for k in range(n_clients):
        cid = first_client + k
        feat, labels, logits = O.synth_client(rows, D, C, seed=seed + cid, signed=signed)
        feat2, _, logits2 = O.synth_client(rows, D, C, seed=seed + 100 + cid, signed=signed)
        zg = torch.randn(rows, C, generator=torch.Generator().manual_seed(seed + 200 + cid)) * 2
        active = [cid % C]
        clients.append(dict(feat=feat, labels=labels, logits=logits, feat2=feat2, logits2=logits2, zg=zg, client_id=cid,
                            active=active, missing=[c for c in range(C) if c not in active], idxs=list(range(rows)),
                            targets_np=labels.numpy().copy()))
    proto = O.synth_prototypes(clients[0]["feat"], clients[0]["labels"])
    return clients, proto


def round_step(ref, clients, proto, state_dicts, weights, C, clean_frac=0.005, noise_frac=0.01, L=0.3, U=0.7):
    """One round over ALL the clients + the server aggregation (main.py:218-234).  Returns timings."""
    args = types.SimpleNamespace(annotation_num=1, n_classes=C)
    t0 = time.perf_counter()
    outs = [client_step(ref, cl, proto, args, clean_frac, noise_frac, L, U) for cl in clients]
    t1 = time.perf_counter()
    w_glob = ref.FedAvg.FedAvg(state_dicts, weights)
    t2 = time.perf_counter()
    n = len(clients)
    class_active = [[k for k in range(n) if c in clients[k]["active"]] for c in range(C)]
    class_negative = [[k for k in range(n) if c in clients[k]["missing"]] for c in range(C)]
    tao = ref.FedAvg.FedAvg_tao([o["t"] for o in outs], weights[:n], class_negative)
    with np.errstate(all="ignore"):
        protos = ref.FedAvg.FedAvg_proto([o["proto"] for o in outs], weights[:n], class_active)
    t3 = time.perf_counter()
    return dict(t_clients=t1 - t0, t_fedavg=t2 - t1, t_tail=t3 - t2, outs=outs, w_glob=w_glob, tao=tao, proto=protos)
