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
The CNN forward/backward stays stock PyTorch / cuDNN (out of the hot path's scope).  Two SURVEY §8f
rows are wired in when the model allows it:
  * f1  a model that exposes `tagging(x, table, sim_out, col0)` (fedmlp_b200.FusedTail: backbone
        `features` + `classifier`) runs the tagging pass through the fused pool+score kernel: the batch's
        similarities land in their columns of the tagger's [C, N] matrix and the [N, D] feature matrix of
        :1026-1049 is never materialised;
  * f2  a model whose parameters were flattened (`flat.flatten_module_`, what `run_fedmlp_rounds` does once
        per client) is stepped by `optim.FlatAdam` — one launch per step — and FedAvg reads its flat buffer
        in place; the three deep copies per client per round of main.py:181,196,219 become one flat copy.

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
from .flat import FlatStateDict, flatten_module_
from .losses import fedmlp_stage1_loss, fedmlp_stage2_loss
from .optim import FlatAdam
from .pooling import build_sim_table
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
        self.sim_mode = getattr(args, "sim_mode", "pair")     # "pair" = the reference's op order; "folded" halves the FMAs
        self.traindata_idx, self.idxss = [], []
        self.last = {}     # tensors of the last call (used by the tests)
        logging.info(f"---> Client{client_id}, each class num: {self.class_num_list}, total num: {len(self.idxs)}")

    # ------------------------------------------------------------------------------------
    def _loader(self, batch_size, shuffle):
        return DataLoader(self.view, batch_size=batch_size, shuffle=shuffle, num_workers=self.num_workers)

    @torch.no_grad()
    def _extract(self, net, shuffle=False, want_features=True, table=None):
        """features [N, D] and logits [N, C] of the local data, rows in idxs order (:1026-1049, :1223-1227).
        shuffle=True walks the data like the reference's `self.ldr_train` (a shuffled loader, :1026); rows are
        scattered back by position, so only the consumption of the global RNG depends on it.
        table (a SimTable) + a model with `.tagging`: the fused pool+score kernel fills self.tagger.sim batch by
        batch and no [N, D] matrix exists (returns (None, logits))."""
        net.eval()
        fused = table is not None and hasattr(net, "tagging")
        feat = logits = None
        N = len(self.idxs)
        for samples, _, pos in self._loader(self.args.batch_size * 4, shuffle=shuffle and not fused):
            x = samples["image_aug_1"].to(self.device, non_blocking=True)
            if fused:           # batches are consecutive row ranges (shuffle is off): columns col0 .. col0 + B
                f, z = net.tagging(x, table, self.tagger.sim, int(pos[0]))
            else:
                f, z = net(x)
            if logits is None:
                logits = torch.empty(N, z.shape[1], dtype=torch.float32, device=self.device)
                if want_features and not fused:
                    feat = torch.empty(N, f.shape[1], dtype=torch.float32, device=self.device)
            pos = pos.to(self.device)
            if feat is not None:
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
    def _make_optimizer(self, net):
        """torch.optim.Adam(lr, betas=(0.9, 0.999), weight_decay=5e-4) re-created every round (:912-913,
        :1149-1150).  A flattened model gets the fused one-launch FlatAdam with the same hyper-parameters; its
        moment buffers are kept across rounds and only zeroed (fresh optimizer state, no allocation)."""
        flat = getattr(net, "_fmlp_flat", None)
        if flat is None:
            return torch.optim.Adam(net.parameters(), lr=self.lr, betas=(0.9, 0.999), weight_decay=5e-4)
        opt = getattr(net, "_fmlp_adam", None)
        if opt is None:
            opt = net._fmlp_adam = FlatAdam(net, lr=self.lr, betas=(0.9, 0.999), weight_decay=5e-4, flat=flat)
        opt.reset(lr=self.lr)
        return opt

    def train_FedMLP(self, rnd, tao, Prototype, writer1, negetive_class_list, active_class_list_client_i, net,
                     glob_model=None):
        """glob_model (extension): the frozen incoming global model, when the caller already holds it (the
        round loop passes netglob itself); default = deepcopy(net) like the reference (:908, :1018)."""
        args = self.args
        stage1 = rnd < args.rounds_FedMLP_stage1
        if glob_model is None:
            glob_model = deepcopy(net)
            for p in glob_model.parameters():
                p.requires_grad_(False)
        glob_model.eval()
        self.optimizer = self._make_optimizer(net)
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
        Prototype = torch.as_tensor(Prototype).to(self.device, dtype=torch.float32)
        if hasattr(net, "tagging"):
            # f1: relu + avg-pool + similarity in one pass over every batch's feature map; no [N, D] matrix
            table = build_sim_table(Prototype, neg, mode=self.sim_mode)
            feat, _ = self._extract(net, want_features=False, table=table)
            self.tagger.select(args.clean_threshold, args.noise_threshold)
        else:
            feat, _ = self._extract(net, shuffle=True)                           # features of the incoming global model
            self.tagger.step(feat, Prototype, args.clean_threshold, args.noise_threshold, mode=self.sim_mode)
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
def run_fedmlp_rounds(args, netglob, trainers, dict_len, rounds, writer1=None, on_round_end=None, flat=True):
    """The FedMLP branch of main.py's round loop (:106-237) with its typos repaired
    (`'FeMLP'`/`train_FeMLP`).  trainers: list of LocalUpdate; netglob: the global
    model (on args.device).  Returns (tao, Prototype, per-round mean client loss).

    flat=True: instead of `deepcopy(netglob)` per client per round, `deepcopy(w_local)` and
    `load_state_dict(deepcopy(w_glob))` (main.py:181,196,219), every client owns ONE model whose parameters
    are views of a flat buffer (flatten_module_, made once); a round starts with one flat copy of the global
    buffer into it, FlatAdam steps it with one launch, FedAvg reads the K flat buffers in place and the
    result is one flat copy back into netglob (int64 BatchNorm counters truncate on load like
    load_state_dict does).  flat=False keeps the reference's copies."""
    n = len(trainers)
    active_class_list, negetive_class_list = [], []
    class_active_client_list, class_negative_client_list = [], []
    tao, Prototype, history = [0] * args.n_classes, [], []
    client_nets = None
    if flat:
        netglob.to(args.device)
        glob_flat = flatten_module_(netglob)
        client_nets = []
        for _ in trainers:
            net = deepcopy(netglob).to(args.device)          # once, not per round
            net._fmlp_flat = flatten_module_(net)
            client_nets.append(net)
    for rnd in range(rounds):
        w_locals, loss_locals, taos, Prototypes = [], [], [], []
        for i, local in enumerate(trainers):
            if flat:
                net = client_nets[i]
                net._fmlp_flat.flat_f32.copy_(glob_flat.flat_f32)            # "deepcopy(netglob)": one flat copy
                if glob_flat.flat_i64 is not None:
                    net._fmlp_flat.flat_i64.copy_(glob_flat.flat_i64)
                teacher = netglob                                             # frozen during the round
            else:
                net = deepcopy(netglob).to(args.device)
                teacher = None
            if rnd < args.rounds_FedMLP_stage1 - 1:
                ret = local.train_FedMLP(rnd, tao, Prototype, writer1, None, None, net, glob_model=teacher)
            else:
                ret = local.train_FedMLP(rnd, tao, Prototype, writer1, negetive_class_list[i], active_class_list[i], net,
                                         glob_model=teacher)
                taos.append(deepcopy(ret[6]))
                Prototypes.append(ret[7].clone())
            if rnd == 0:
                active_class_list.append(ret[5])
                negetive_class_list.append(ret[4])
            w_locals.append(net._fmlp_flat if flat else ret[0])
            loss_locals.append(ret[1])
        if rnd == 0:                                                              # main.py:200-210
            for c in range(args.n_classes):
                class_active_client_list.append([j for j in range(n) if c in active_class_list[j]])
                class_negative_client_list.append([j for j in range(n) if c in negetive_class_list[j]])
        assert len(w_locals) == len(dict_len) == n                                # main.py:212
        w_glob = FedAvg(w_locals, dict_len)                                       # :218/221
        if flat:
            lay = glob_flat.layout
            glob_flat.flat_f32.copy_(w_glob.flat_f32[:lay.n_f32])                 # load_state_dict: one flat copy
            if glob_flat.flat_i64 is not None:
                glob_flat.flat_i64.copy_(w_glob.flat_f32[lay.n_f32:lay.n_f32 + lay.n_i64])   # float32 -> int64 truncation
        else:
            netglob.load_state_dict(w_glob)
        if rnd >= args.rounds_FedMLP_stage1 - 1:
            tao = FedAvg_tao(taos, dict_len, class_negative_client_list)          # :223
            Prototype = FedAvg_proto(Prototypes, dict_len, class_active_client_list)   # :231/234 (lam = 1)
        history.append(float(np.mean(loss_locals)))
        if on_round_end is not None:
            on_round_end(rnd, netglob, tao, Prototype)
    return tao, Prototype, history
