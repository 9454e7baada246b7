"""One FedMLP round hot path for all simulated clients resident on one GPU.

The reference simulates clients one after the other in a Python loop (main.py:135) and every
client runs tag -> train -> prototypes on its own (utils/local_training.py:1006-1256), then the
server averages (main.py:218-234).  On a B200 a whole shard of clients lives on the device at
once: their rows are stored back to back in [N_total, D] matrices ("segments"), and each stage
of the round is ONE launch over all of them:

    tag         fmlp_tag_sim_f32 + fmlp_tag_select + fmlp_mask_fill       (K3 / K3b / K3c)
    loss        fmlp_loss_stage2_f32 per client over its N_k rows          (K4)
    prototypes  fmlp_proto_build_f32                                       (K2)
    FedAvg      fmlp_fedavg_flat_f32 over the clients' flat parameter buffers (K1)

This module is host-side orchestration only; bench.py and the multi-GPU driver (dist.py) use it.
The CNN forward/backward (cuDNN) sits between these stages in real training and is not part of
the hot path measured here.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import torch

from .fedavg import fedavg_flat_buffers
from . import _cabi as cabi
from .losses import launch_stage2
from .prototypes import PrototypeResult, build_prototypes
from .tagging import TagBatch


@dataclass
class RoundResult:
    counts: torch.Tensor            # [S, C, 4] n_clean, n_noise, m, k
    sel: torch.Tensor               # [S, C, 2, cap] selected rows (rank order)
    losses: torch.Tensor            # [S] stage-2 loss per client over its rows
    dz: torch.Tensor                # [N_total, C] gradient w.r.t. the logits
    protos: PrototypeResult         # per-client prototypes / counts / t counts
    global_flat: torch.Tensor       # [P] aggregated parameters
    events: dict = field(default_factory=dict)


class ClientShard:
    """The clients of one rank: tagging state + the batched round hot path."""

    def __init__(self, sizes, n_classes, active_classes, device=None, clean_frac=0.005, noise_frac=0.01,
                 L=0.3, U=0.7, sim_mode="pair", dataset_idx=None):
        self.sizes = [int(n) for n in sizes]
        self.S = len(self.sizes)
        self.C = int(n_classes)
        self.seg_rows = [0]
        for n in self.sizes:
            self.seg_rows.append(self.seg_rows[-1] + n)
        self.N = self.seg_rows[-1]
        self.active = [list(a) for a in active_classes]
        self.missing = [[c for c in range(self.C) if c not in a] for a in self.active]
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.clean_frac, self.noise_frac, self.L, self.U = clean_frac, noise_frac, L, U
        self.sim_mode = sim_mode
        self.tagger = TagBatch(self.seg_rows, self.C, self.active, self.missing, dataset_idx=dataset_idx,
                               device=self.device)

    def round_hot_path(self, feat_tag, proto_glob, logits, logits_glob, labels, feat_proto, logits_proto,
                       client_flats, weights, timers=None, fedavg_out=None, divide=True, divisor=None) -> RoundResult:
        """feat_tag [N, D]: features of the incoming global model (tagging, :1026-1049);
        logits / logits_glob [N, C]: student / frozen-global logits for the loss (:1178-1188);
        feat_proto / logits_proto: features and logits of the locally trained model (:1223-1239);
        client_flats: S flat parameter buffers [P]; weights: S client weights (dict_len)."""
        ev = {}

        def mark(name):
            if timers is not None:
                e = torch.cuda.Event(enable_timing=True)
                e.record(torch.cuda.current_stream(self.device))
                ev[name] = e

        mark("start")
        self.tagger.similarity(feat_tag, proto_glob, self.sim_mode)
        mark("sim")
        counts, sel, cap = self.tagger.select(self.clean_frac, self.noise_frac)
        y, distill, sup = self.tagger.fill(labels)
        mark("select_fill")
        losses = torch.empty(self.S, dtype=torch.float32, device=self.device)
        dz = torch.empty_like(logits)
        for s in range(self.S):
            r0, r1 = self.seg_rows[s], self.seg_rows[s + 1]
            launch_stage2(logits[r0:r1], logits_glob[r0:r1], y[r0:r1], distill[r0:r1], cabi.LOSS2_SUP,
                          losses[s:s + 1], dz[r0:r1])
        mark("loss")
        protos = build_prototypes(feat_proto, labels, logits_proto, self.active, self.missing, self.L, self.U,
                                  guard_empty=True, seg_rows=self.seg_rows)
        mark("proto")
        glob = fedavg_flat_buffers(client_flats, weights, out=fedavg_out, divide=divide, divisor=divisor)
        mark("fedavg")
        return RoundResult(counts, sel, losses, dz, protos, glob, ev)
