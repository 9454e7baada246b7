"""Multi-GPU FedMLP rounds: clients are sharded over the ranks, only aggregation crosses GPUs.

The reference is single-process / single-GPU and simulates its clients sequentially
(main.py:32,135); its "communication" is a Python list of state_dicts (main.py:130,196).  Here
one process drives one GPU (torchrun), each rank owns a contiguous block of the clients and runs
tag / loss / prototypes for them with NO collective.  The only exchange step per round is the
server aggregation (main.py:218-234):

    FedAvg        rank-local K_r-way weighted partial sum (fedavg.cu, weights pre-normalised by the
                  global total) -> ONE all-reduce(sum) of the flat fp32 parameter buffer over
                  NCCL / NVLink -> every rank holds the identical global model.  int64 BatchNorm
                  counters are all-reduced as exact int64 weighted sums and divided afterwards
                  (bit-identical to the reference); the fp32 parameters differ from the
                  reference's client-sequential fold only by summation association (<= 1e-6 rel).
    FedAvg_proto  per-class weighted means over the clients that annotate the class
    FedAvg_tao    per-class weighted means of the difficulty statistics (float64)

`local_reduce` is injectable so the distributed algebra can be exercised on CPU with the gloo
backend in tests (the CUDA kernel is the default and the only product path).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from . import _cabi as cabi
from .fedavg import FedAvg_proto, _is_integral, fedavg_flat_buffers
from .flat import FlatStateDict, layout_of


def shard_clients(n_clients: int, world: int, rank: int) -> range:
    """Contiguous block of client ids owned by `rank` (sizes differ by at most one)."""
    base, extra = divmod(n_clients, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def _world(group):
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


def _cuda_local_reduce(bufs, weights, out):
    """sum_i bufs[i] * weights[i] on the GPU (fmlp_fedavg_flat_f32 without the final divide)."""
    return fedavg_flat_buffers(bufs, weights, out=out, divide=False)


def global_weight_sum(local_weights, group=None, device=None) -> float:
    """sum of dict_len over all ranks (float64 all-reduce of one scalar)."""
    t = torch.tensor([float(sum(local_weights))], dtype=torch.float64, device=device or "cpu")
    if _world(group) > 1:
        dist.all_reduce(t, group=group)
    return float(t.item())


def fedavg_flat_distributed(local_bufs, local_weights, total_weight=None, out=None, group=None,
                            local_reduce=None):
    """Global weighted mean of every client's flat fp32 buffer; each rank passes only its own
    clients.  Returns the [P] mean (identical on every rank).  total_weight: sum of all clients'
    weights over all ranks (computed with one extra scalar all-reduce when omitted)."""
    if len(local_bufs) == 0:
        raise ValueError("every rank needs at least one client")
    dev = local_bufs[0].device
    if total_weight is None:
        total_weight = global_weight_sum(local_weights, group, dev if dev.type == "cuda" else None)
    w = [float(x) / float(total_weight) for x in local_weights]   # float64 division, rounded once to fp32 by the kernel ABI
    reduce_fn = local_reduce or _cuda_local_reduce
    if out is None:
        out = torch.empty_like(local_bufs[0])
    partial = reduce_fn(local_bufs, w, out)
    if _world(group) > 1:
        dist.all_reduce(partial, group=group)
    return partial


class FusedFedAvgAllReduce:
    """Local K-way weighted fold + chunk-pipelined two-shot all-reduce over NVLink peer memory in ONE
    kernel per rank (fedavg_allreduce.cu).  Symmetric-memory buffers are allocated and exchanged once;
    every call is a single cooperative launch on the current stream.  Collective: all ranks call it with
    the same P.  Needs NVLink/P2P between the ranks' GPUs (torch symmetric memory)."""

    FLAG_WORDS = 512      # FMLP_AR_FLAG_WORDS
    MAX_CHUNKS = 16       # FMLP_AR_MAX_CHUNKS

    def __init__(self, P: int, group=None, device=None, n_chunks=None):
        import os

        import torch.distributed._symmetric_memory as symm

        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        if self.world > 8:
            raise ValueError("FusedFedAvgAllReduce supports up to 8 ranks (one NVSwitch domain)")
        if P % 4:
            raise ValueError("P must be a multiple of 4 (flat buffers are 16-byte padded)")
        self.P = int(P)
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        # pipeline chunks: the all-gather of chunk c-1 overlaps the fold + reduce-scatter of chunk c+1
        if n_chunks is None:
            n_chunks = int(os.environ.get("FMLP_AR_CHUNKS", "2"))   # r01 sweep: 2 is best at 2 and 8 GPUs
        self.n_chunks, self.L = self.chunk_layout(self.P, self.world, n_chunks)
        n = self.world * self.L
        self.stage = symm.empty(n, dtype=torch.float32, device=self.device)
        self.result = symm.empty(n, dtype=torch.float32, device=self.device)
        self.flags = symm.empty(self.FLAG_WORDS, dtype=torch.int32, device=self.device)
        self.stage.zero_(); self.result.zero_(); self.flags.zero_()
        hs = symm.rendezvous(self.stage, self.group)
        hr = symm.rendezvous(self.result, self.group)
        hf = symm.rendezvous(self.flags, self.group)
        self._handles = (hs, hr, hf)
        self.stage_ptrs = cabi.ptr_array(list(hs.buffer_ptrs))
        self.result_ptrs = cabi.ptr_array(list(hr.buffer_ptrs))
        self.flag_ptrs = cabi.ptr_array(list(hf.buffer_ptrs))
        self.epoch_dev = torch.zeros(1, dtype=torch.int32, device=self.device)   # per-rank call counter
        torch.cuda.synchronize(self.device)
        dist.barrier(self.group)          # every rank's flags are zero before anyone signals

    @classmethod
    def chunk_layout(cls, P: int, world: int, n_chunks: int):
        """(n_chunks, L): the parameter vector is cut into n_chunks chunks of `world` slices of L / n_chunks
        floats (a multiple of 4, so every slice is 16-byte aligned); L = floats per rank over all chunks."""
        n_chunks = max(1, min(cls.MAX_CHUNKS, int(n_chunks)))
        per_slice = (P + world * n_chunks - 1) // (world * n_chunks)
        return n_chunks, (per_slice + 3) // 4 * 4 * n_chunks

    def __call__(self, local_bufs, weights_normalised):
        """Returns the [P] global weighted mean (a view of this rank's symmetric result buffer,
        overwritten by the next call)."""
        K = len(local_bufs)
        if not 1 <= K <= cabi.MAX_CLIENTS:
            raise ValueError(f"1..{cabi.MAX_CLIENTS} clients per rank")
        cabi.require_cuda(*local_bufs)
        for b in local_bufs:
            if self.P and (b.numel() != self.P or b.dtype != torch.float32 or not b.is_contiguous()):
                raise ValueError("client buffers must be contiguous float32 of length P")
        with torch.cuda.device(self.device):
            cabi.check(cabi.lib().fmlp_fedavg_allreduce_f32(
                cabi.ptr_array([b.data_ptr() for b in local_bufs]), cabi.f32_array(weights_normalised), K, self.P,
                self.stage_ptrs, self.result_ptrs, self.flag_ptrs, self.L, self.n_chunks, self.rank, self.world,
                self.epoch_dev.data_ptr(),
                cabi.stream_ptr(self.device)), "fmlp_fedavg_allreduce_f32")
        return self.result[:self.P]


class QueuedAggregation:
    """Round-2 aggregation exchange (csrc/fedavg_allreduce_q.cu): local K-way weighted fold + all-reduce of
    [P parameters | T tail floats | M fp64 scalars] in ONE non-cooperative work-queue kernel per rank.  The
    reduction goes through the NVSwitch (multimem.ld_reduce / multimem.st) when torch's symmetric memory
    exposes a multicast mapping, otherwise by peer loads in rank order + peer stores.  Symmetric buffers are
    allocated and exchanged once; every call is a single ordinary launch on the current stream (CUDA-graph
    capturable, overlaps kernels of other streams).  Collective: all ranks call it with the same shapes."""

    FLAG_WORDS = 512      # FMLP_AR_FLAG_WORDS

    def __init__(self, P: int, T: int = 0, M: int = 0, group=None, device=None, n_chunks=None, max_ctas=None,
                 use_multicast=None, fold_iters=0, red_iters=0):
        import os

        import torch.distributed._symmetric_memory as symm

        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        if self.world > 8:
            raise ValueError("QueuedAggregation supports up to 8 ranks (one NVSwitch domain)")
        if P % 4 or T % 4:
            raise ValueError("P and T must be multiples of 4 (16-byte padded buffers)")
        self.P, self.T, self.M = int(P), int(T), int(M)
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        lib = cabi.lib()
        n = int(lib.fmlp_fedavg_allreduce_q_buffer_floats(self.P, self.T, self.M))
        if n_chunks is None:     # r02 sweeps (profiles/r02_arq_sweep_*gpu.jsonl): 4 chunks at 2 GPUs, 8 at 8
            n_chunks = int(os.environ.get("FMLP_ARQ_CHUNKS", "8" if self.world >= 4 else "4"))
        self.n_chunks = max(1, min(16, int(n_chunks)))
        self.max_ctas = int(os.environ.get("FMLP_ARQ_CTAS", "0")) if max_ctas is None else int(max_ctas)
        self.fold_iters, self.red_iters = int(fold_iters), int(red_iters)
        self.partial = symm.empty(n, dtype=torch.float32, device=self.device)
        self.result = symm.empty(n, dtype=torch.float32, device=self.device)
        self.flags = symm.empty(self.FLAG_WORDS, dtype=torch.int32, device=self.device)
        self.partial.zero_(); self.result.zero_(); self.flags.zero_()
        hp = symm.rendezvous(self.partial, self.group)
        hr = symm.rendezvous(self.result, self.group)
        hf = symm.rendezvous(self.flags, self.group)
        self._handles = (hp, hr, hf)

        def peer_ptrs(h, t):
            # buffer_ptrs are the bases of the allocation blocks; the tensor may sit at an offset inside its block
            off = t.data_ptr() - int(h.buffer_ptrs[self.rank])
            return [int(b) + off for b in h.buffer_ptrs], off

        pp, off_p = peer_ptrs(hp, self.partial)
        rp, off_r = peer_ptrs(hr, self.result)
        fp, _ = peer_ptrs(hf, self.flags)
        self.partial_ptrs, self.result_ptrs, self.flag_ptrs = cabi.ptr_array(pp), cabi.ptr_array(rp), cabi.ptr_array(fp)
        if use_multicast is None:
            use_multicast = os.environ.get("FMLP_ARQ_MULTICAST", "1") != "0"
        self.mc_partial = self.mc_result = 0
        if use_multicast and self.world > 1:
            try:
                mp_, mr_ = int(hp.multicast_ptr), int(hr.multicast_ptr)
                if mp_ and mr_:
                    self.mc_partial, self.mc_result = mp_ + off_p, mr_ + off_r
            except Exception:      # no multicast support on this system: peer-to-peer path
                self.mc_partial = self.mc_result = 0
        # every rank must take the same path (the flags protocol is the same, the data path is not)
        ok = torch.tensor([1 if self.mc_partial else 0], dtype=torch.int32, device=self.device)
        if self.world > 1:
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
        if int(ok.item()) == 0:
            self.mc_partial = self.mc_result = 0
        self.nvls = bool(self.mc_partial)
        self.epoch_dev = torch.zeros(1, dtype=torch.int32, device=self.device)   # per-rank call counter
        torch.cuda.synchronize(self.device)
        dist.barrier(self.group)          # every rank's flags are zero before anyone signals

    @property
    def path(self) -> str:
        return "nvls multimem" if self.nvls else "peer-to-peer pull/push"

    def __call__(self, local_bufs, weights_normalised, tail_bufs=None, tail_f64=None):
        """local_bufs: K contiguous fp32 [P] tensors; tail_bufs: K fp32 [T] tensors (T > 0); tail_f64: device
        float64 [M] (M > 0).  Returns (params [P], tail [T], f64 [M]) — views of this rank's symmetric result
        buffer, overwritten by the next call."""
        K = len(local_bufs)
        if not 1 <= K <= cabi.MAX_CLIENTS:
            raise ValueError(f"1..{cabi.MAX_CLIENTS} clients per rank")
        cabi.require_cuda(*local_bufs)
        for b in local_bufs:
            if self.P and (b.numel() != self.P or b.dtype != torch.float32 or not b.is_contiguous()):
                raise ValueError("client buffers must be contiguous float32 of length P")
        tails = None
        if self.T:
            if tail_bufs is None or len(tail_bufs) != K:
                raise ValueError("one tail vector per client")
            for b in tail_bufs:
                if b.numel() != self.T or b.dtype != torch.float32 or not b.is_contiguous() or not b.is_cuda:
                    raise ValueError("tail vectors must be contiguous float32 CUDA tensors of length T")
            tails = cabi.ptr_array([b.data_ptr() for b in tail_bufs])
        if self.M and (tail_f64 is None or tail_f64.numel() != self.M or tail_f64.dtype != torch.float64 or not tail_f64.is_cuda):
            raise ValueError("tail_f64 must be a float64 CUDA tensor of length M")
        with torch.cuda.device(self.device):
            cabi.check(cabi.lib().fmlp_fedavg_allreduce_q_f32(
                cabi.ptr_array([b.data_ptr() for b in local_bufs]), tails, cabi.f32_array(weights_normalised), K,
                self.P, self.T, tail_f64.data_ptr() if self.M else None, self.M,
                self.partial_ptrs, self.result_ptrs, self.flag_ptrs, self.mc_partial or None, self.mc_result or None,
                self.n_chunks, self.rank, self.world, self.epoch_dev.data_ptr(), self.max_ctas, self.fold_iters,
                self.red_iters, cabi.stream_ptr(self.device)), "fmlp_fedavg_allreduce_q_f32")
        r = self.result
        f64 = r[self.P + self.T:self.P + self.T + 2 * self.M].view(torch.float64) if self.M else None
        return r[:self.P], (r[self.P:self.P + self.T] if self.T else None), f64


class FedMLPAggregation:
    """The whole server aggregation of a FedMLP round (main.py:218-234) for the clients of this rank, across
    ranks: the FedAvg parameters (utils/FedAvg.py:7-14), the aggregated prototypes (FedAvg_proto, :72-93), tao
    (FedAvg_tao, :51-70, float64) and the int64 BatchNorm counters as float32 (FedAvg.py:13), identical on every
    rank.  split=False: ONE exchange carries [parameters | prototype sums | fp64 tail] (tail pack -> queue kernel
    -> finalize).  split=True: the parameters have their own exchange, which only needs the clients' weights and
    can start at the very beginning of the round on its own stream, and the small tails (2C*D floats + 3C+J
    doubles) follow the prototype pass in a second, single-chunk launch of the same kernel."""

    def __init__(self, P: int, C: int, D: int, J: int = 0, group=None, device=None, split=False, params_impl="queue", **kw):
        """params_impl (split=True only): "queue" = the work-queue kernel, "push" = the round-1 peer-store kernel
        (FusedFedAvgAllReduce: cooperative launch, reduce-scatter by posted peer stores)."""
        self.C, self.D, self.J = int(C), int(D), int(J)
        self.T = 2 * self.C * self.D
        self.M = 3 * self.C + self.J
        self.split = bool(split)
        self.params_impl = params_impl if self.split else "queue"
        if self.split and self.params_impl == "push":
            self.exchange = FusedFedAvgAllReduce(P, group=group, device=device)
            self.tails_exchange = QueuedAggregation(0, self.T, self.M, group=group, device=device, n_chunks=1, max_ctas=8,
                                                    use_multicast=kw.get("use_multicast"))
        elif self.split:
            self.exchange = QueuedAggregation(P, 0, 0, group=group, device=device, **kw)
            self.tails_exchange = QueuedAggregation(0, self.T, self.M, group=group, device=device, n_chunks=1, max_ctas=8,
                                                    use_multicast=kw.get("use_multicast"))
        else:
            self.exchange = QueuedAggregation(P, self.T, self.M, group=group, device=device, **kw)
            self.tails_exchange = self.exchange
        dev = self.exchange.device
        self.tail_local = torch.zeros(self.M, dtype=torch.float64, device=dev)
        self.proto = torch.empty(2 * self.C, self.D, dtype=torch.float32, device=dev)
        self.tao = torch.empty(self.C, dtype=torch.float64, device=dev)
        self.counters = torch.empty(max(self.J, 1), dtype=torch.float32, device=dev)

    def _pack(self, tcnt, weights, rows, active, missing, counters, st):
        K = len(weights)
        cabi.check(cabi.lib().fmlp_agg_tail_pack_f64(
            tcnt.data_ptr() if tcnt is not None else None, K, self.C, cabi.f64_array(weights), cabi.i64_array(rows),
            cabi.u32_array([cabi.class_mask(a) for a in active]), cabi.u32_array([cabi.class_mask(m) for m in missing]),
            cabi.ptr_array([c.data_ptr() for c in counters]) if self.J else None, self.J,
            self.tail_local.data_ptr(), st), "fmlp_agg_tail_pack_f64")

    def _finalize(self, psum, f64, total_weight, st):
        cabi.check(cabi.lib().fmlp_agg_finalize_f32(psum.data_ptr(), f64.data_ptr(), self.C, self.D, self.J, float(total_weight),
                                                    self.proto.data_ptr(), self.tao.data_ptr(),
                                                    self.counters.data_ptr() if self.J else None, st), "fmlp_agg_finalize_f32")
        return self.proto, self.tao, (self.counters[:self.J] if self.J else None)

    def aggregate_params(self, client_flats, weights, total_weight):
        """split=True: the parameter exchange alone (current stream)."""
        wn = [float(w) / float(total_weight) for w in weights]
        if self.params_impl == "push":
            return self.exchange(client_flats, wn)
        params, _, _ = self.exchange(client_flats, wn)
        return params

    def aggregate_tails(self, client_protos, tcnt, weights, rows, active, missing, total_weight, counters=None):
        """split=True: prototypes / tao / counters (current stream; after the prototype pass)."""
        if not self.split:
            raise RuntimeError("aggregate_tails needs split=True (the combined exchange carries the tails itself)")
        ex = self.tails_exchange
        dev = ex.device
        wn = [float(w) / float(total_weight) for w in weights]
        tails = [p.reshape(-1) for p in client_protos]
        with torch.cuda.device(dev):
            st = cabi.stream_ptr(dev)
            self._pack(tcnt, weights, rows, active, missing, counters, st)
            # P == 0: the fp32 vector is the tail alone (the kernel still wants K valid source pointers)
            _, psum, f64 = ex(tails, wn, tail_bufs=tails, tail_f64=self.tail_local)
            return self._finalize(psum, f64, total_weight, st)

    def __call__(self, client_flats, client_protos, tcnt, weights, rows, active, missing, total_weight, counters=None):
        """client_flats: K fp32 [P]; client_protos: K fp32 [2C, D] (rows of classes the client does not annotate
        are zero, as utils/local_training.py:973-1002 leaves them); tcnt: int32 [K, C] confident counts;
        weights / rows: K client weights (dict_len) and row counts; active / missing: K class lists;
        total_weight: sum of weights over ALL ranks; counters: K int64 [J] tensors or None."""
        if self.split:
            params = self.aggregate_params(client_flats, weights, total_weight)
            proto, tao, cnt = self.aggregate_tails(client_protos, tcnt, weights, rows, active, missing, total_weight, counters)
            return params, proto, tao, cnt
        ex = self.exchange
        dev = ex.device
        wn = [float(w) / float(total_weight) for w in weights]
        with torch.cuda.device(dev):
            st = cabi.stream_ptr(dev)
            self._pack(tcnt, weights, rows, active, missing, counters, st)
            params, psum, f64 = ex(client_flats, wn, tail_bufs=[p.reshape(-1) for p in client_protos], tail_f64=self.tail_local)
            proto, tao, cnt = self._finalize(psum, f64, total_weight, st)
        return params, proto, tao, cnt


def FedAvg_distributed(w_local, dict_len_local, group=None, total_weight=None, local_reduce=None):
    """Distributed drop-in for FedAvg(w, dict_len) (reference utils/FedAvg.py:7-14): every rank
    passes the state_dicts and weights of ITS clients; all ranks get the same averaged dict
    (same keys/order; int64 entries become float32 like the reference)."""
    if len(w_local) == 0:
        raise ValueError("every rank needs at least one client")
    first = next(iter(w_local[0].values()))
    dev = first.device
    layout = layout_of(w_local[0])
    flats = []
    for sd in w_local:
        if layout_of(sd) is not layout:
            raise KeyError("FedAvg_distributed: clients have different state_dict layouts")
        if isinstance(sd, FlatStateDict) and not getattr(sd, "ints_as_float", False):
            flats.append(sd)
        else:
            flats.append(FlatStateDict.from_state_dict(sd, device=dev))   # pack once (copy)
    if total_weight is None:
        total_weight = global_weight_sum(dict_len_local, group, dev if dev.type == "cuda" else None)
    out = FlatStateDict.empty(layout, dev, ints_as_float=True)
    if layout.n_f32:
        fedavg_flat_distributed([f.flat_f32 for f in flats], dict_len_local, total_weight,
                                out=out.flat_f32[:layout.n_f32], group=group, local_reduce=local_reduce)
    if layout.n_i64:
        integral = all(_is_integral(x) for x in dict_len_local)
        if integral:
            acc = torch.zeros(layout.n_i64, dtype=torch.int64, device=dev)
            for f, n in zip(flats, dict_len_local):
                acc += f.flat_i64 * int(n)                      # exact int64, like w[k] * dict_len[i]
            if _world(group) > 1:
                dist.all_reduce(acc, group=group)
            out.flat_f32[layout.n_f32:] = acc.to(torch.float32) / float(total_weight)
        else:
            acc = torch.zeros(layout.n_i64, dtype=torch.float32, device=dev)
            for f, n in zip(flats, dict_len_local):
                acc += f.flat_i64.to(torch.float32) * float(n)
            if _world(group) > 1:
                dist.all_reduce(acc, group=group)
            out.flat_f32[layout.n_f32:] = acc / float(total_weight)
    return out


def FedAvg_proto_distributed(protos_local, weight_local, class_active_local, n_classes, group=None,
                             local_proto_avg=None):
    """Distributed FedAvg_proto (reference utils/FedAvg.py:72-93).  protos_local: this rank's
    client prototypes [2C, D]; class_active_local[c] = LOCAL client positions annotating class c.
    Each rank forms its local per-class weighted mean, scales it by its share of the class
    weight and one all-reduce adds the shares.  Classes nobody annotates give NaN rows, like the
    reference's 0/0."""
    avg_fn = local_proto_avg or FedAvg_proto
    local = avg_fn(protos_local, weight_local, class_active_local)          # [2C, D]; NaN rows where empty
    dev = local.device
    wc = torch.zeros(n_classes, dtype=torch.float64, device=dev)
    for c, clients in enumerate(class_active_local):
        wc[c] = float(sum(weight_local[i] for i in clients))
    total = wc.clone()
    if _world(group) > 1:
        dist.all_reduce(total, group=group)
    share = torch.where(total > 0, wc / total.clamp_min(1e-300), torch.zeros_like(wc)).to(torch.float32)
    rows = share.repeat_interleave(2).unsqueeze(1)                          # [2C, 1]
    contrib = torch.where(rows > 0, local * rows, torch.zeros_like(local))  # drop this rank's NaN rows
    if _world(group) > 1:
        dist.all_reduce(contrib, group=group)
    nobody = (total == 0).repeat_interleave(2).unsqueeze(1)
    return torch.where(nobody, torch.full_like(contrib, float("nan")), contrib)


def FedAvg_tao_distributed(t_local, weight_local, class_client_local, n_classes, group=None):
    """Distributed FedAvg_tao (reference utils/FedAvg.py:51-70, float64): class_client_local[c] =
    LOCAL client positions for which class c is missing; a class nobody misses gets 1.0."""
    num = np.zeros(n_classes, dtype=np.float64)
    den = np.zeros(n_classes, dtype=np.float64)
    for c, clients in enumerate(class_client_local):
        for i in clients:
            num[c] += t_local[i][c] * float(weight_local[i])
            den[c] += float(weight_local[i])
    packed = torch.from_numpy(np.concatenate([num, den]))
    if _world(group) > 1:
        backend = dist.get_backend(group)
        if backend == "nccl":
            packed = packed.cuda()
        dist.all_reduce(packed, group=group)
        packed = packed.cpu()
    num, den = packed[:n_classes].numpy(), packed[n_classes:].numpy()
    return np.where(den > 0, num / np.where(den > 0, den, 1.0), 1.0)
