"""Drop-in `LocalUpdate.train_FedMLP` (reference utils/local_training.py:27-56, :904-1256) and the
FedMLP round loop of main.py (:85-106,130-135,178-237) on top of the fedmlp_b200 kernels.

Same call surface as the reference:

    local = LocalUpdate(args, client_id, dataset, idxs, class_pos_idx, class_neg_idx,
                        active_class_list=[i], student=..., teacher_neg=..., teacher_act=...)
    ret = local.train_FedMLP(rnd, tao, Prototype, writer1, negetive_class_list,
                             active_class_list_client_i, net)
    # rnd <  stage1-1 : (state_dict, mean_loss, _, _, neg_list, act_list)
    # rnd >= stage1-1 : (state_dict, mean_loss, _, _, neg_list, act_list, t, proto)

The model contract is the reference's: `net(images) -> (feature [B, D], logits [B, C])`; the
dataset yields the reference's sample dicts (`image_aug_1`, `image_aug_2`, `target`, `index`).
What changes underneath: labels / masks / tagging state live on the GPU (no per-sample Python
membership scans, no `.cpu()` / `.item()` / `.tolist()` syncs inside the loops), the losses are the
fused kernels, prototypes and similarities are one pass over a preallocated [N, D] feature buffer.
The CNN forward/backward and Adam stay stock PyTorch / cuDNN (out of the hot path's scope).

Deliberate deviations from the letter of the reference (none changes values):
  * rows are kept in `idxs` order; the reference's shuffled-loader order only affects ties between
    exactly equal similarities;
  * the returned state_dict stays on the GPU (the reference moves the net to the CPU at :1251,
    which forces FedAvg onto the host); `FedAvg` accepts either.
"""
from __future__ import annotations

import logging
from copy import deepcopy

import numpy as np
import torch
from torch.utils.data import DataLoader, Dataset

from .fedavg import FedAvg, FedAvg_proto, FedAvg_tao
from .losses import fedmlp_stage1_loss, fedmlp_stage2_loss
from .prototypes import build_prototypes
from .tagging import TagBatch


class _IndexedView(Dataset):
    """(sample dict, dataset index) pairs of one client, in `idxs` order (DatasetSplit :1328-1356
    without the per-item label mutation — labels are served from the device-resident matrices)."""

    def __init__(self, dataset, idxs):
        self.dataset, self.idxs = dataset, list(idxs)

    def __len__(self):
        return len(self.idxs)

    def __getitem__(self, item):
        s = self.dataset[self.idxs[item]]
        return {"image_aug_1": s["image_aug_1"], "image_aug_2": s["image_aug_2"]}, self.idxs[item], item


class LocalUpdate(object):
    def __init__(self, args, client_id, dataset, idxs, class_pos_idx, class_neg_idx, active_class_list=None,
                 student=None, teacher_neg=None, teacher_act=None, dataset_test=None, num_workers=0):
        self.args, self.client_id, self.dataset = args, client_id, dataset
        self.idxs = [int(i) for i in idxs]
        self.student, self.teacher_neg, self.teacher_act, self.dataset_test = student, teacher_neg, teacher_act, dataset_test
        self.class_pos_idx, self.class_neg_idx = class_pos_idx, class_neg_idx
        self.device = torch.device(getattr(args, "device", "cuda"))
        if self.device.type != "cuda":
            raise RuntimeError("fedmlp_b200.LocalUpdate needs a CUDA device (no CPU fallback)")
        C = args.n_classes
        self.active_class_list = list(active_class_list) if active_class_list is not None else \
            sorted(np.random.choice(C, args.annotation_num, replace=False).tolist())
        self.negative_class_list = [c for c in range(C) if c not in self.active_class_list]
        # labels as the client sees them in stage 1: positives of non-annotated classes listed in
        # class_neg_idx are hidden (DatasetSplit.__getitem__ :1347-1351)
        targets = np.asarray(dataset.targets, dtype=np.float32)[self.idxs].copy()
        self.targets_true = torch.from_numpy(targets.copy()).to(self.device)
        idx_arr = np.asarray(self.idxs)
        for c in self.negative_class_list:
            hidden = np.isin(idx_arr, np.asarray(class_neg_idx[c]))
            targets[hidden, c] = 0
        self.labels = torch.from_numpy(targets).to(self.device)            # [N, C]
        # reference bookkeeping kept as attributes (:39-44)
        self.class_num_list = np.asarray(dataset.targets, dtype=np.float64)[self.idxs].sum(0).tolist()
        self.loss_w = [len(self.idxs) / i if i else float("inf") for i in self.class_num_list]
        self.view = _IndexedView(dataset, self.idxs)
        self.num_workers = num_workers
        self.epoch, self.iter_num, self.lr = 0, 0, args.base_lr
        self.tagger = None
        self.traindata_idx, self.idxss = [], []
        self.last = {}     # tensors of the last call (used by the tests)
        logging.info(f"---> Client{client_id}, each class num: {self.class_num_list}, total num: {len(self.idxs)}")

    # ------------------------------------------------------------------------------------
    def _loader(self, batch_size, shuffle):
        return DataLoader(self.view, batch_size=batch_size, shuffle=shuffle, num_workers=self.num_workers)

    @torch.no_grad()
    def _extract(self, net):
        """features [N, D] and logits [N, C] of the local data, rows in idxs order (:1026-1049, :1223-1227)."""
        net.eval()
        feat = logits = None
        for samples, _, pos in self._loader(self.args.batch_size * 4, shuffle=False):
            f, z = net(samples["image_aug_1"].to(self.device, non_blocking=True))
            if feat is None:
                feat = torch.empty(len(self.idxs), f.shape[1], dtype=torch.float32, device=self.device)
                logits = torch.empty(len(self.idxs), z.shape[1], dtype=torch.float32, device=self.device)
            pos = pos.to(self.device)
            feat[pos] = f.float()
            logits[pos] = z.float()
        return feat, logits

    def _prototypes(self, net, guard_empty):
        feat, logits = self._extract(net)
        res = build_prototypes(feat, self.labels, logits, self.active_class_list, self.negative_class_list,
                               self.args.L, self.args.U, guard_empty=guard_empty)
        self.last.update(proto_feat=feat, proto_logits=logits)
        return res.t()[0], res.proto[0].cpu()

    # ------------------------------------------------------------------------------------
    def train_FedMLP(self, rnd, tao, Prototype, writer1, negetive_class_list, active_class_list_client_i, net):
        args = self.args
        stage1 = rnd < args.rounds_FedMLP_stage1
        glob_model = deepcopy(net)
        glob_model.eval()
        for p in glob_model.parameters():
            p.requires_grad_(False)
        self.optimizer = torch.optim.Adam(net.parameters(), lr=self.lr, betas=(0.9, 0.999), weight_decay=5e-4)
        act, neg = self.active_class_list, self.negative_class_list
        epoch_loss, step_losses = [], []
        self.last = {"steps": []}

        if stage1:                                                              # :907-1004
            for c in neg:
                self.class_num_list[c] = 0
            net.train()
            for _ in range(args.local_ep):
                batch_loss = []
                for samples, _, pos in self._loader(args.batch_size, shuffle=True):
                    x1 = samples["image_aug_1"].to(self.device, non_blocking=True)
                    x2 = samples["image_aug_2"].to(self.device, non_blocking=True)
                    labels = self.labels[pos.to(self.device)]
                    _, z1 = net(x1)
                    _, z2 = net(x2)
                    with torch.no_grad():
                        _, z3 = glob_model(x1)
                        _, z4 = glob_model(x2)
                    loss = fedmlp_stage1_loss(z1, z2, z3, z4, labels, act, neg, args.batch_size)
                    self.optimizer.zero_grad()
                    loss.backward()
                    self.optimizer.step()
                    batch_loss.append(loss.detach())
                    self.last["steps"].append(dict(z1=z1.detach(), z2=z2.detach(), z3=z3, z4=z4, y=labels, loss=loss.detach()))
                    self.iter_num += 1
                self.epoch += 1
                epoch_loss.append(torch.stack(batch_loss).mean().item())       # one sync per epoch
            if rnd == args.rounds_FedMLP_stage1 - 1:                            # first t and prototypes (:971-1002)
                t, proto = self._prototypes(net, guard_empty=False)
                return net.state_dict(), float(np.mean(epoch_loss)), None, None, neg, act, t, proto
            return net.state_dict(), float(np.mean(epoch_loss)), None, None, neg, act

        # ---------------------------------------------------------------- stage 2 (:1006-1256)
        neg = list(negetive_class_list) if negetive_class_list is not None else neg
        if self.tagger is None or rnd == args.rounds_FedMLP_stage1:
            self.tagger = TagBatch([0, len(self.idxs)], args.n_classes, [act], [neg],
                                   dataset_idx=torch.tensor(self.idxs), device=self.device)
        feat, _ = self._extract(net)                                             # features of the incoming global model
        Prototype = torch.as_tensor(Prototype).to(self.device, dtype=torch.float32)
        self.tagger.step(feat, Prototype, args.clean_threshold, args.noise_threshold)
        self.last.update(tag_feat=feat, prototype=Prototype)
        y_all, distill_all, _ = self.tagger.fill(self.targets_true)             # DatasetSplit_pseudo (:1456-1477)
        noise_counts = self.tagger.class_num_noise(0)
        for i, c in enumerate(neg):                                              # :1117-1120
            self.class_num_list[c] = noise_counts[i]
        net.train()
        for _ in range(args.local_ep):
            batch_loss = []
            for samples, _, pos in self._loader(args.batch_size, shuffle=True):
                x1 = samples["image_aug_1"].to(self.device, non_blocking=True)
                rows = pos.to(self.device)
                labels, distill = y_all[rows], distill_all[rows]
                _, z = net(x1)
                with torch.no_grad():
                    _, zg = glob_model(x1)
                loss = fedmlp_stage2_loss(z, zg, labels, distill)                # :1188
                self.optimizer.zero_grad()
                loss.backward()
                self.optimizer.step()
                batch_loss.append(loss.detach())
                self.last["steps"].append(dict(z=z.detach(), zg=zg, y=labels, distill=distill, loss=loss.detach()))
                self.iter_num += 1
            self.epoch += 1
            epoch_loss.append(torch.stack(batch_loss).mean().item())
        self.traindata_idx = self.tagger.traindata_idx(0)
        self.idxss = self.tagger.remaining(0)
        t, proto = self._prototypes(net, guard_empty=True)                        # :1208-1249
        self.optimizer.zero_grad()
        return net.state_dict(), float(np.mean(epoch_loss)), None, None, neg, act, t, proto


# ---------------------------------------------------------------------------------------- round loop
def run_fedmlp_rounds(args, netglob, trainers, dict_len, rounds, writer1=None, on_round_end=None):
    """The FedMLP branch of main.py's round loop (:106-237) with its typos repaired
    (`'FeMLP'`/`train_FeMLP`).  trainers: list of LocalUpdate; netglob: the global
    model (on args.device).  Returns (tao, Prototype, per-round mean client loss)."""
    n = len(trainers)
    active_class_list, negetive_class_list = [], []
    class_active_client_list, class_negative_client_list = [], []
    tao, Prototype, history = [0] * args.n_classes, [], []
    for rnd in range(rounds):
        w_locals, loss_locals, taos, Prototypes = [], [], [], []
        for i, local in enumerate(trainers):
            net = deepcopy(netglob).to(args.device)
            if rnd < args.rounds_FedMLP_stage1 - 1:
                ret = local.train_FedMLP(rnd, tao, Prototype, writer1, None, None, net)
            else:
                ret = local.train_FedMLP(rnd, tao, Prototype, writer1, negetive_class_list[i], active_class_list[i], net)
                taos.append(deepcopy(ret[6]))
                Prototypes.append(ret[7].clone())
            if rnd == 0:
                active_class_list.append(ret[5])
                negetive_class_list.append(ret[4])
            w_locals.append(ret[0])
            loss_locals.append(ret[1])
        if rnd == 0:                                                              # main.py:200-210
            for c in range(args.n_classes):
                class_active_client_list.append([j for j in range(n) if c in active_class_list[j]])
                class_negative_client_list.append([j for j in range(n) if c in negetive_class_list[j]])
        assert len(w_locals) == len(dict_len) == n                                # main.py:212
        w_glob = FedAvg(w_locals, dict_len)                                       # :218/221
        netglob.load_state_dict(w_glob)
        if rnd >= args.rounds_FedMLP_stage1 - 1:
            tao = FedAvg_tao(taos, dict_len, class_negative_client_list)          # :223
            Prototype = FedAvg_proto(Prototypes, dict_len, class_active_client_list)   # :231/234 (lam = 1)
        history.append(float(np.mean(loss_locals)))
        if on_round_end is not None:
            on_round_end(rnd, netglob, tao, Prototype)
    return tao, Prototype, history
