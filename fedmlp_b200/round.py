"""One FedMLP round hot path for all simulated clients resident on one GPU.

The reference simulates clients one after the other in a Python loop (main.py:135) and every
client runs tag -> train -> prototypes on its own (utils/local_training.py:1006-1256), then the
server averages (main.py:218-234).  On a B200 a whole shard of clients lives on the device at
once: their rows are stored back to back in [N_total, D] matrices ("segments"), and each stage
of the round is ONE launch over all of them:

    tag         fmlp_tag_sim_f32 + fmlp_tag_select + fmlp_mask_fill       (K3 / K3b / K3c)
    loss        fmlp_loss_stage2_seg_f32: every client's rows, own denominators (K4)
    prototypes  fmlp_proto_build_f32                                       (K2)
    FedAvg      fmlp_fedavg_flat_f32 over the clients' flat parameter buffers (K1)

This module is host-side orchestration only; bench.py and the multi-GPU driver (dist.py) use it.
Memory is planned once (`_Plan`): every output and workspace is a persistent device buffer and
every host-side argument array is pre-built, so a round is seven C calls and no allocation.
The CNN forward/backward (cuDNN) sits between these stages in real training and is not part of
the hot path measured here.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass, field

import torch

from . import _cabi as cabi
from .prototypes import PrototypeResult
from .tagging import SIM_MODES, TagBatch


@dataclass
class RoundResult:
    counts: torch.Tensor            # [S, C, 4] n_clean, n_noise, m, k
    sel: torch.Tensor               # [S, C, 2, cap] selected rows (rank order)
    losses: torch.Tensor            # [S] stage-2 loss per client over its rows
    dz: torch.Tensor                # [N_total, C] gradient w.r.t. the logits
    protos: PrototypeResult         # per-client prototypes / counts / t counts
    global_flat: torch.Tensor       # [P] aggregated parameters
    events: dict = field(default_factory=dict)
    proto_glob: torch.Tensor = None  # [2C, D] aggregated prototypes (FedAvg_proto), when the tails are aggregated
    tao: torch.Tensor = None         # [C] float64 (FedAvg_tao)
    counters: torch.Tensor = None    # [J] float32 aggregated int64 BatchNorm counters


class _Plan:
    """Static buffers + pre-built ctypes arguments for one (D, P) configuration."""

    def __init__(self, shard: "ClientShard", D: int, P: int):
        lib = cabi.lib()
        dev, S, C, N = shard.device, shard.S, shard.C, shard.N
        self.D, self.P = D, P
        f32, i32 = torch.float32, torch.int32
        max_rows = max(shard.sizes, default=0)
        cap = max(1, int(math.floor(max(shard.clean_frac, shard.noise_frac, 0.0) * max_rows)) + 1)
        self.cap = min(cap, max(max_rows, 1))
        self.counts = torch.zeros(S, C, 4, dtype=i32, device=dev)
        self.remaining = torch.zeros(S, C, dtype=i32, device=dev)
        self.sel = torch.full((S, C, 2, self.cap), -1, dtype=i32, device=dev)
        self.y = torch.empty(N, C, dtype=f32, device=dev)
        self.distill = torch.empty(N, C, dtype=f32, device=dev)
        self.sup = torch.empty(N, C, dtype=f32, device=dev)
        self.losses = torch.empty(S, dtype=f32, device=dev)
        self.dz = torch.empty(N, C, dtype=f32, device=dev)
        self.proto = torch.empty(S, 2 * C, D, dtype=f32, device=dev)
        self.cnt = torch.empty(S, 2 * C, dtype=i32, device=dev)
        self.tcnt = torch.zeros(S, C, dtype=i32, device=dev)
        self.glob = torch.empty(P, dtype=f32, device=dev)
        self.one = torch.ones(2, dtype=f32, device=dev)
        # aggregation tails (FedAvg_proto / FedAvg_tao / int64 counters, main.py:218-234)
        self.proto_glob = torch.empty(2 * C, D, dtype=f32, device=dev)
        self.tao = torch.empty(C, dtype=torch.float64, device=dev)
        self.tail = torch.zeros(3 * C + 1024, dtype=torch.float64, device=dev)
        self.counters = torch.empty(1024, dtype=f32, device=dev)
        self.class_active = cabi.u64_array([sum(1 << k for k in range(S) if c in shard.active[k]) for c in range(C)])
        self.sizes = cabi.i64_array(shard.sizes)
        with torch.cuda.device(dev):
            self.ws_sim = torch.empty(max(lib.fmlp_tag_sim_ws_bytes(C, D), 256), dtype=torch.uint8, device=dev)
            self.ws_select = torch.empty(max(lib.fmlp_tag_select_ws_bytes(S, C, self.cap), 256), dtype=torch.uint8, device=dev)
            self.ws_loss = torch.zeros(max(lib.fmlp_loss_ws_bytes(N, C), 256), dtype=torch.uint8, device=dev)   # arrival counter starts at 0
            self.ws_proto = torch.empty(max(lib.fmlp_proto_ws_bytes(N, D, C, S), 256), dtype=torch.uint8, device=dev)
        self.rows = cabi.i64_array(shard.seg_rows)
        self.active = cabi.u32_array([cabi.class_mask(a) for a in shard.active])
        self.missing = cabi.u32_array([cabi.class_mask(m) for m in shard.missing])


class ClientShard:
    """The clients of one rank: tagging state + the batched round hot path."""

    # traindata_idx bookkeeping costs two small device copies per round; benchmarks switch it off
    keep_history = True
    # shared-memory pad (KB) of the prototype CTAs in the single-GPU two-stream round, None = leave the knob alone
    proto_pad_smem_kb = 40
    # single-GPU two-stream round: "concurrent" = similarity and prototype kernels start together; "sim_first" /
    # "proto_first" = one of the two streaming kernels runs alone and the other starts behind it (same work, same
    # results; which is fastest depends on the shape: bench.py measures the three and keeps the best)
    schedule = "concurrent"

    def __init__(self, sizes, n_classes, active_classes, device=None, clean_frac=0.005, noise_frac=0.01,
                 L=0.3, U=0.7, sim_mode="folded", dataset_idx=None):
        """sim_mode: "folded" (default here: one dot product per class against
        q_c = P0/|P0| - P1/|P1|, runs at the HBM ceiling) or "pair" (two cosines then subtract, the
        reference's op order; FMA-co-limited for >= 4 missing classes).  They differ by O(1e-7),
        inside the north_star waiver; both are parity-tested on the recorded reference flow."""
        self.sizes = [int(n) for n in sizes]
        self.S = len(self.sizes)
        if self.S > cabi.MAX_SEGMENTS:
            raise ValueError(f"a ClientShard holds at most {cabi.MAX_SEGMENTS} clients; use several shards")
        self.C = int(n_classes)
        self.seg_rows = [0]
        for n in self.sizes:
            self.seg_rows.append(self.seg_rows[-1] + n)
        self.N = self.seg_rows[-1]
        self.active = [list(a) for a in active_classes]
        self.missing = [[c for c in range(self.C) if c not in a] for a in self.active]
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.clean_frac, self.noise_frac, self.L, self.U = float(clean_frac), float(noise_frac), float(L), float(U)
        self.sim_mode = sim_mode
        self.loss_variant = cabi.LOSS2_SUP      # cabi.LOSS2_SUP_DIS: the commented variant of :1187 (BCE + teacher MSE)
        self.tagger = TagBatch(self.seg_rows, self.C, self.active, self.missing, dataset_idx=dataset_idx,
                               device=self.device)
        self._plan = None
        self._tail_stream = None

    def plan(self, D, P) -> _Plan:
        if self._plan is None or self._plan.D != D or self._plan.P != P:
            self._plan = _Plan(self, D, P)
        return self._plan

    def round_hot_path(self, feat_tag, proto_glob, logits, logits_glob, labels, feat_proto, logits_proto,
                       client_flats, weights, timers=None, fedavg_out=None, divide=True, divisor=None,
                       side_stream=None, after_aggregate=None, aggregate_fn=None, counters=None,
                       aggregate_tails=False, agg_stream=None, params_fn=None, tails_fn=None, only=None) -> RoundResult:
        """feat_tag [N, D]: features of the incoming global model (tagging, :1026-1049);
        logits / logits_glob [N, C]: student / frozen-global logits for the loss (:1178-1188);
        feat_proto / logits_proto: features and logits of the locally trained model (:1223-1239);
        client_flats: S flat parameter buffers [P]; weights: S client weights (dict_len).
        All inputs are contiguous fp32 CUDA tensors on this shard's device.  The returned tensors
        are the shard's persistent buffers (overwritten by the next round).

        The pass is a DAG over its inputs: {sim -> select -> mask fill -> loss} needs the incoming
        global model's features, {prototypes -> aggregation} only the locally trained model.
        side_stream: run prototypes + aggregation on this stream, concurrently with the tagging/loss chain
        (the latency-bound select / fill / loss kernels hide behind the streaming ones; in the live
        loop the same split overlaps them with the cuDNN work around them).  Without a side stream the
        stages run back to back as sim, prototypes, aggregation, select, fill, loss: the small kernels
        follow the aggregation, whose 28 MB of freshly written lines are still being written back.
        after_aggregate(glob): called with the aggregation stream current right after the FedAvg launch —
        the NCCL path issues its all-reduce there.
        aggregate_fn(client_flats, weights, protos) -> [P] tensor (or a tuple (params, proto_glob, tao,
        counters)): replaces the FedAvg launch altogether (the fused fold + all-reduce kernels of dist.py).
        agg_stream (with side_stream): three-way split — the parameter aggregation only needs the clients'
        weights, so it starts at the very beginning of the round on agg_stream, the prototype pass and the small
        tails run on side_stream, the tagging/loss chain on the current stream.
        params_fn(client_flats, weights) -> [P] tensor and tails_fn(protos) -> (proto_glob, tao, counters): the
        multi-GPU exchanges of the split form (dist.FedMLPAggregation(split=True)).
        only: "sim" | "proto" | "fedavg" | "tail" — launch just that stage on the current stream (bench.py times
        each stage's launches back to back this way).
        aggregate_tails / counters: single-GPU aggregation of the small tails of main.py:218-234 next to
        FedAvg — FedAvg_proto (bit-exact kernel), FedAvg_tao (float64) and the int64 BatchNorm counters
        (counters: S int64 tensors of equal length)."""
        N, C, S = self.N, self.C, self.S
        D = feat_tag.shape[1]
        if (tuple(feat_tag.shape) != (N, D) or tuple(feat_proto.shape) != (N, D) or tuple(labels.shape) != (N, C)
                or tuple(logits.shape) != (N, C) or tuple(proto_glob.shape) != (2 * C, D) or len(client_flats) != S):
            raise ValueError("round_hot_path: inconsistent shapes")
        for t in (feat_tag, proto_glob, logits, logits_glob, labels, feat_proto, logits_proto, *client_flats):
            if not t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous():
                raise TypeError("round_hot_path needs contiguous float32 CUDA tensors (no CPU fallback)")
        P = client_flats[0].numel()
        pl = self.plan(D, P)
        lib = cabi.lib()
        check = cabi.check
        tg = self.tagger
        glob = pl.glob if fedavg_out is None else fedavg_out
        if divisor is None:
            divisor = sum(weights)
        ev = {}
        dev = self.device
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev)
            st = stream.cuda_stream

            def mark(name):
                if timers is not None:
                    # timers="external": events that become record nodes when the round is captured
                    # in a CUDA graph (per-stage timing without eager launch gaps)
                    e = (torch.cuda.Event(enable_timing=True, external=True) if timers == "external"
                         else torch.cuda.Event(enable_timing=True))
                    e.record(stream)
                    ev[name] = e

            out = {"glob": glob, "proto_glob": None, "tao": None, "counters": None}
            protos = PrototypeResult(pl.proto, pl.cnt, pl.tcnt, list(self.seg_rows))

            def proto_stage(stream_b):
                sb = stream_b.cuda_stream
                check(lib.fmlp_proto_build_f32(feat_proto.data_ptr(), D, D, labels.data_ptr(), logits_proto.data_ptr(), 0,
                                               C, S, pl.rows, pl.active, pl.missing, self.L, self.U, 1,
                                               pl.proto.data_ptr(), pl.cnt.data_ptr(), pl.tcnt.data_ptr(),
                                               pl.ws_proto.data_ptr(), pl.ws_proto.numel(), sb), "fmlp_proto_build_f32")
                if stream_b is stream:
                    mark("proto")

            def params_stage(stream_b):
                sb = stream_b.cuda_stream
                if params_fn is not None:
                    out["glob"] = params_fn(client_flats, weights)
                    return
                flags = cabi.FEDAVG_DIVIDE if divide else 0
                check(lib.fmlp_fedavg_flat_f32(cabi.ptr_array([b.data_ptr() for b in client_flats]),
                                               cabi.f32_array(weights), S, P, float(divisor), flags, glob.data_ptr(),
                                               sb), "fmlp_fedavg_flat_f32")
                if after_aggregate is not None:
                    after_aggregate(glob)

            def tails_stage(stream_b):
                sb = stream_b.cuda_stream
                if tails_fn is not None:
                    out["proto_glob"], out["tao"], out["counters"] = tails_fn(protos)
                    return
                if not aggregate_tails:
                    return
                J = int(counters[0].numel()) if counters else 0
                if J > 1024:
                    raise ValueError("at most 1024 int64 counters")
                # FedAvg_proto (utils/FedAvg.py:72-93, bit-exact), FedAvg_tao (:51-70, float64) and the int64
                # counters (:9-13) in one launch (three dependent launches until round 2)
                check(lib.fmlp_agg_tails_local_f32(pl.proto.data_ptr(), S, C, D, cabi.f64_array(weights), pl.class_active,
                                                   pl.proto_glob.data_ptr(), pl.tcnt.data_ptr(), pl.sizes, pl.missing,
                                                   cabi.ptr_array([c.data_ptr() for c in counters]) if J else None, J,
                                                   float(divisor), pl.tao.data_ptr(), pl.counters.data_ptr() if J else None,
                                                   sb), "fmlp_agg_tails_local_f32")
                out["proto_glob"], out["tao"] = pl.proto_glob, pl.tao
                out["counters"] = pl.counters[:J] if J else None

            def aggregate_stage(stream_b):
                if aggregate_fn is not None:
                    r = aggregate_fn(client_flats, weights, protos)
                    if isinstance(r, tuple):
                        out["glob"], out["proto_glob"], out["tao"], out["counters"] = r
                    else:
                        out["glob"] = r
                else:
                    params_stage(stream_b)
                    tails_stage(stream_b)
                if stream_b is stream:
                    mark("fedavg")

            def select_fill_loss():
                check(lib.fmlp_tag_select(tg.sim.data_ptr(), tg.sim.shape[1], tg.tag.data_ptr(), tg.tag.shape[1], C, S,
                                          pl.rows, pl.missing, self.clean_frac, self.noise_frac, pl.counts.data_ptr(),
                                          pl.remaining.data_ptr(), pl.sel.data_ptr(), pl.cap, pl.ws_select.data_ptr(),
                                          pl.ws_select.numel(), st), "fmlp_tag_select")
                mark("select_fill")
                # label / mask fill fused into the loss: one ordinary launch behind the selection (round 2)
                check(lib.fmlp_fill_loss_stage2_f32(labels.data_ptr(), tg.tag.data_ptr(), tg.tag.shape[1], logits.data_ptr(),
                                                    logits_glob.data_ptr(), C, S, pl.rows, pl.active, pl.missing,
                                                    self.loss_variant, pl.remaining.data_ptr(), pl.y.data_ptr(),
                                                    pl.distill.data_ptr(), pl.sup.data_ptr(), pl.losses.data_ptr(),
                                                    pl.dz.data_ptr(), pl.ws_loss.data_ptr(), pl.ws_loss.numel(), st),
                      "fmlp_fill_loss_stage2_f32")
                mark("loss")

            if timers == "external":
                # a one-element kernel in front of the first event: the launch latency of the graph itself is
                # then not attributed to the first timed stage
                check(lib.fmlp_scale_f32(pl.one.data_ptr(), 1, pl.one.data_ptr() + 4, st), "fmlp_scale_f32")
            def sim_stage():
                check(lib.fmlp_tag_sim_f32(feat_tag.data_ptr(), D, D, proto_glob.data_ptr(), C, S, pl.rows, pl.missing,
                                           tg.sim.data_ptr(), tg.sim.shape[1], SIM_MODES[self.sim_mode], pl.ws_sim.data_ptr(),
                                           pl.ws_sim.numel(), st), "fmlp_tag_sim_f32")

            if only is not None:
                {"sim": sim_stage, "proto": lambda: proto_stage(stream), "fedavg": lambda: aggregate_stage(stream),
                 "tail": select_fill_loss}[only]()
                return RoundResult(pl.counts, pl.sel, pl.losses, pl.dz, protos, out["glob"], ev)
            sim_launched = False
            schedule = os.environ.get("FMLP_ROUND_SCHEDULE") or self.schedule
            if schedule not in ("concurrent", "sim_first", "proto_first"):
                raise ValueError(f"unknown round schedule {schedule!r}")
            tail_stream = None
            if side_stream is not None and agg_stream is None and aggregate_tails and tails_fn is None and aggregate_fn is None:
                if self._tail_stream is None:
                    self._tail_stream = torch.cuda.Stream(device=dev)
                tail_stream = self._tail_stream
            mark("start")
            split3 = side_stream is not None and agg_stream is not None and aggregate_fn is None
            if split3:
                agg_stream.wait_stream(stream)
                with torch.cuda.stream(agg_stream):
                    params_stage(agg_stream)
                side_stream.wait_stream(stream)
                with torch.cuda.stream(side_stream):
                    proto_stage(side_stream)
                    tails_stage(side_stream)
            elif side_stream is not None:
                if schedule == "sim_first":
                    # the similarity kernel runs alone; the prototype pass starts behind it, next to select / fill+loss
                    sim_stage()
                    mark("sim")
                    sim_launched = True
                side_stream.wait_stream(stream)
                with torch.cuda.stream(side_stream):
                    # Single-GPU two-stream round: the prototype pass runs next to the similarity kernel.  Its CTAs
                    # request 40 KB of (unused) shared memory each, so that at most one of them shares an SM with a
                    # similarity CTA (measured: round 0.131 -> 0.124 ms, profiles/r02_exp_coresidency.txt).
                    pad_prev = None
                    if tail_stream is not None and self.proto_pad_smem_kb is not None:
                        pad_prev = lib.fmlp_get_tuning(cabi.TUNE_PROTO_PAD_SMEM_KB)
                        if pad_prev < 0:      # an explicit setting (or the environment variable) wins
                            if os.environ.get("FMLP_PROTO_PAD_SMEM_KB") is None:
                                lib.fmlp_set_tuning(cabi.TUNE_PROTO_PAD_SMEM_KB, int(self.proto_pad_smem_kb))
                            else:
                                pad_prev = None
                        else:
                            pad_prev = None
                    proto_stage(side_stream)
                    if pad_prev is not None:
                        lib.fmlp_set_tuning(cabi.TUNE_PROTO_PAD_SMEM_KB, pad_prev)
                    if schedule == "proto_first":
                        # the prototype pass runs alone; the similarity kernel starts behind it, next to FedAvg
                        proto_done = torch.cuda.Event()
                        proto_done.record(side_stream)
                        stream.wait_event(proto_done)
                    if aggregate_fn is None and tail_stream is not None:
                        # the small tails only need the prototypes: they leave the chain here and run next to the
                        # parameter aggregation instead of behind it (latency-bound, 7 us at the end of the round)
                        tail_stream.wait_stream(side_stream)
                        with torch.cuda.stream(tail_stream):
                            tails_stage(tail_stream)
                        params_stage(side_stream)
                    else:
                        aggregate_stage(side_stream)
            if not sim_launched:
                sim_stage()
                mark("sim")
            if side_stream is not None:
                select_fill_loss()
                stream.wait_stream(side_stream)
                if split3:
                    stream.wait_stream(agg_stream)
                elif aggregate_fn is None and tail_stream is not None:
                    stream.wait_stream(tail_stream)
            else:
                proto_stage(stream)
                aggregate_stage(stream)
                select_fill_loss()
        if self.keep_history:
            # keep the lazily materialised host lists of the tagger in sync with this round's picks
            tg._history.append((pl.counts.clone(), pl.sel.clone(), pl.cap))
        tg._lists = None
        return RoundResult(pl.counts, pl.sel, pl.losses, pl.dz, protos, out["glob"], ev,
                           proto_glob=out["proto_glob"], tao=out["tao"], counters=out["counters"])
