"""Server aggregation — drop-in for the reference's utils/FedAvg.py.

Same call surface: `FedAvg(w, dict_len)`, `Fed_w(w, weight)`, `FedAvg_proto(Prototypes, weight,
class_active_client_list)`, `FedAvg_tao(t, weight, class_active_client_list=None)`
(reference: utils/FedAvg.py:7-14, :16-23, :72-93, :51-70; call sites main.py:218-234).

`FedAvg` returns an OrderedDict with the reference's keys, order and result dtypes (fp32 stays
fp32, int64 BatchNorm counters become float32 — the true-division quirk) on the
device of the inputs.  The arithmetic runs in libfedmlp_b200 (fedavg.cu): one launch over the
flat parameter buffers when the clients are `FlatStateDict`s (or views laid out like one), else
one multi-tensor launch that reads the K x T scattered tensors in place.
"""
from __future__ import annotations

import numbers
from collections import OrderedDict

import numpy as np
import torch

from . import _cabi as cabi
from .flat import FlatLayout, FlatStateDict, flat_view_of, layout_of

_plan_cache: dict = {}


def _is_integral(x) -> bool:
    return isinstance(x, (numbers.Integral, np.integer)) and not isinstance(x, bool)


def _multi_plan(layout: FlatLayout, device):
    """Device-resident chunk tables of the multi-tensor kernel for one layout (cached)."""
    key = (layout, str(device))
    plan = _plan_cache.get(key)
    if plan is not None:
        return plan
    f_idx, i_idx = layout.float_index, layout.int_index
    chunk_tensor, chunk_start, tensor_chunk0 = [], [], []
    for t, i in enumerate(f_idx):
        n = layout.numels[i]
        tensor_chunk0.append(len(chunk_tensor))
        for s in range(0, n, cabi.FEDAVG_CHUNK):
            chunk_tensor.append(t)
            chunk_start.append(s)
    tensor_chunk0.append(len(chunk_tensor))
    elem_tensor, elem_index = [], []
    for t, i in enumerate(i_idx):
        for e in range(layout.numels[i]):
            elem_tensor.append(t)
            elem_index.append(e)
    plan = dict(
        f_idx=f_idx, i_idx=i_idx,
        numel=torch.tensor([layout.numels[i] for i in f_idx], dtype=torch.int64, device=device),
        chunk_tensor=torch.tensor(chunk_tensor, dtype=torch.int32, device=device),
        chunk_start=torch.tensor(chunk_start, dtype=torch.int64, device=device),
        n_chunks=len(chunk_tensor),
        tensor_chunk0=torch.tensor(tensor_chunk0, dtype=torch.int64, device=device),
        elem_tensor=torch.tensor(elem_tensor, dtype=torch.int32, device=device),
        elem_index=torch.tensor(elem_index, dtype=torch.int64, device=device),
        n_elems=len(elem_tensor),
        f_off=np.array([layout.offsets[i] for i in f_idx], dtype=np.int64),
        i_off=np.array([layout.offsets[i] for i in i_idx], dtype=np.int64),
    )
    _plan_cache[key] = plan
    return plan


def _check_same_keys(w):
    lay0 = getattr(w[0], "layout", None)
    keys = None
    for i in range(1, len(w)):
        if lay0 is not None and getattr(w[i], "layout", None) is lay0:
            continue        # FlatStateDicts of one layout: same keys by construction
        if keys is None:
            keys = list(w[0].keys())
        if list(w[i].keys()) != keys:
            raise KeyError(f"FedAvg: client {i} has different state_dict keys than client 0")


def fedavg_flat_buffers(bufs, weights, out=None, divisor=None, divide=True):
    """K-way weighted fold of K flat fp32 CUDA buffers of equal length (the bench / sweep entry
    point; also what FedAvg uses underneath).  Any K: folds in groups of MAX_CLIENTS, in order."""
    K = len(bufs)
    if K == 0:
        raise ValueError("FedAvg of zero clients")
    cabi.require_cuda(*bufs)
    P = bufs[0].numel()
    dev = bufs[0].device
    if out is None:
        out = torch.empty(P, dtype=torch.float32, device=dev)
    if divisor is None:
        divisor = sum(weights)
    lib = cabi.lib()
    with torch.cuda.device(dev):
        st = cabi.stream_ptr(dev)
        for g0 in range(0, K, cabi.MAX_CLIENTS):
            g1 = min(K, g0 + cabi.MAX_CLIENTS)
            flags = 0
            if g0 > 0:
                flags |= cabi.FEDAVG_ACCUMULATE
            if g1 == K and divide:
                flags |= cabi.FEDAVG_DIVIDE
            srcs = cabi.ptr_array([b.data_ptr() for b in bufs[g0:g1]])
            ws = cabi.f32_array(weights[g0:g1])
            cabi.check(lib.fmlp_fedavg_flat_f32(srcs, ws, g1 - g0, P, float(divisor), flags,
                                                out.data_ptr(), st), "fmlp_fedavg_flat_f32")
    return out


def _fedavg_i64_flat(ptrs, weights, J, divisor, integral, out_ptr, dev, div_flag=cabi.FEDAVG_DIVIDE, tensors=None):
    """int64 BatchNorm counters (utils/FedAvg.py:9-13).  Integer weights keep the fold in int64 until the true
    division; the kernel folds up to 64 clients per launch.  More clients with integer weights: the exact int64
    partial (the reference's own `w[k] * dict_len[i]` adds, J scalars) is formed first and the kernel only does
    the int64 -> float32 divide — a float32 `out` cannot carry an int64 partial between launches."""
    lib = cabi.lib()
    K = len(ptrs)
    st = cabi.stream_ptr(dev)
    if K > cabi.MAX_CLIENTS and integral:
        if tensors is None:
            raise ValueError("int64 FedAvg over more than 64 clients needs the counter tensors")
        acc = torch.zeros(J, dtype=torch.int64, device=dev)
        for t, n in zip(tensors, weights):
            acc += t.reshape(-1) * int(n)
        cabi.check(lib.fmlp_fedavg_flat_i64(cabi.ptr_array([acc.data_ptr()]), cabi.f64_array([1.0]), 1, J, float(divisor), 1,
                                            div_flag, out_ptr, st), "fmlp_fedavg_flat_i64")
        return acc      # caller keeps it alive until the stream has run
    for g0 in range(0, K, cabi.MAX_CLIENTS):
        g1 = min(K, g0 + cabi.MAX_CLIENTS)
        flags = (cabi.FEDAVG_ACCUMULATE if g0 > 0 else 0) | (div_flag if g1 == K else 0)
        cabi.check(lib.fmlp_fedavg_flat_i64(cabi.ptr_array(ptrs[g0:g1]), cabi.f64_array(weights[g0:g1]), g1 - g0, J,
                                            float(divisor), 1 if integral else 0, flags, out_ptr, st),
                   "fmlp_fedavg_flat_i64")
    return None


_last_layout = None


def _fedavg_cuda(w, dict_len, divide=True):
    """divide=False: plain weighted sum (DaAgg, utils/FedNoRo.py:98-103)."""
    global _last_layout
    layout = _last_layout = layout_of(w[0], like=_last_layout)     # rounds repeat the same model: try the last layout first
    dev = next(iter(w[0].values())).device
    K = len(w)
    integral = all(_is_integral(x) for x in dict_len)
    divisor = sum(dict_len) if divide else 1.0
    div_flag = cabi.FEDAVG_DIVIDE if divide else 0
    out = FlatStateDict.empty(layout, dev, ints_as_float=True, lazy=True)
    lib = cabi.lib()
    with torch.cuda.device(dev):
        st = cabi.stream_ptr(dev)
        # per-client layout once (a signature pass over the 727 entries; FlatStateDicts know theirs), then the flat-view
        # probe with that layout — the probe leaves at the first tensor that is not where a flat buffer would have it
        lays = [layout] + [layout_of(sd, like=layout) for sd in w[1:]]
        views = [flat_view_of(sd, lay) if lay is layout else None for sd, lay in zip(w, lays)]
        if all(v is not None for v in views):
            # ---- flat path: K pointers, one streaming launch ---------------------------
            if layout.n_f32:
                bufs = [v[0] for v in views]
                for g0 in range(0, K, cabi.MAX_CLIENTS):
                    g1 = min(K, g0 + cabi.MAX_CLIENTS)
                    flags = (cabi.FEDAVG_ACCUMULATE if g0 > 0 else 0) | (div_flag if g1 == K else 0)
                    cabi.check(lib.fmlp_fedavg_flat_f32(cabi.ptr_array(bufs[g0:g1]), cabi.f32_array(dict_len[g0:g1]),
                                                        g1 - g0, layout.n_f32, float(divisor), flags,
                                                        out.flat_f32.data_ptr(), st), "fmlp_fedavg_flat_f32")
            if layout.n_i64:
                ints = None
                if K > cabi.MAX_CLIENTS and integral:
                    ints = [sd.flat_i64 if isinstance(sd, FlatStateDict) and sd.flat_i64 is not None else
                            torch.cat([v.reshape(-1) for i, v in enumerate(sd.values()) if layout.is_int[i]]) for sd in w]
                out._keepalive = _fedavg_i64_flat([v[1] for v in views], list(dict_len), layout.n_i64, divisor, integral,
                                                  out.flat_f32.data_ptr() + 4 * layout.n_f32, dev, div_flag, tensors=ints)
            return out
        # ---- multi-tensor path: read the scattered tensors in place ---------------------
        if K > cabi.MAX_CLIENTS:
            # the pointer table of the multi-tensor kernel holds 64 clients: pack once and take the flat path
            # (any K, folded in groups of 64 in the reference's client order)
            packed = [sd if (isinstance(sd, FlatStateDict) and not getattr(sd, "ints_as_float", False))
                      else FlatStateDict.from_state_dict(sd, device=dev) for sd in w]
            return _fedavg_cuda(packed, dict_len, divide)
        plan = _multi_plan(layout, dev)
        f_idx, i_idx = plan["f_idx"], plan["i_idx"]
        Tf, Ti = len(f_idx), len(i_idx)
        n_all = len(layout.keys)
        if any(lay is not layout for lay in lays):     # the kernels index every client by client 0's layout
            raise KeyError("FedAvg: the clients' state_dicts differ in keys, shapes or dtypes")
        tensors = [v for sd in w for v in sd.values()]
        if not all(map(torch.Tensor.is_contiguous, tensors)):
            raise ValueError("FedAvg: non-contiguous state_dict tensor")
        # [K, T] pointers of every client tensor in one pass (the only per-call Python work that scales with the
        # 727 x K tensors); the device table is re-uploaded only when a pointer changed since the last call with
        # this layout, through a pinned staging buffer, without a host synchronisation
        ptrs = np.fromiter(map(torch.Tensor.data_ptr, tensors), dtype=np.int64, count=K * n_all).reshape(K, n_all)
        cache = plan.setdefault("table_cache", {})
        ent = cache.get(K)
        if ent is None:
            n_words = Tf * K + Tf + Ti * K + Ti
            ent = cache[K] = dict(ptrs=None, pinned=torch.empty(n_words, dtype=torch.int64, pin_memory=True),
                                  dev=torch.empty(n_words, dtype=torch.int64, device=dev), event=None)
        table_dev = ent["dev"]
        o = Tf * K + Tf
        same_src = ent["ptrs"] is not None and np.array_equal(ent["ptrs"], ptrs)
        if ent["event"] is not None:
            ent["event"].synchronize()          # the previous upload has left the pinned buffer
        table = ent["pinned"].numpy()
        if not same_src:
            if Tf:
                table[:Tf * K] = ptrs[:, f_idx].T.reshape(-1)
            if Ti:
                table[o:o + Ti * K] = ptrs[:, i_idx].T.reshape(-1)
            ent["ptrs"] = ptrs
        if Tf:
            table[Tf * K:Tf * K + Tf] = out.flat_f32.data_ptr() + 4 * plan["f_off"]
        if Ti:
            table[o + Ti * K:] = out.flat_f32.data_ptr() + 4 * (layout.n_f32 + plan["i_off"])
        table_dev.copy_(ent["pinned"], non_blocking=True)     # 50 KB; stream-ordered before the launches below
        ent["event"] = torch.cuda.Event()
        ent["event"].record(torch.cuda.current_stream(dev))
        base = table_dev.data_ptr()
        if Tf:
            cabi.check(lib.fmlp_fedavg_multi_f32(base, base + 8 * Tf * K, plan["numel"].data_ptr(),
                                                 plan["chunk_tensor"].data_ptr(), plan["chunk_start"].data_ptr(),
                                                 plan["n_chunks"], Tf, cabi.f32_array(dict_len), K, float(divisor),
                                                 div_flag, st), "fmlp_fedavg_multi_f32")
        if Ti:
            cabi.check(lib.fmlp_fedavg_multi_i64(base + 8 * o, base + 8 * (o + Ti * K), plan["elem_tensor"].data_ptr(),
                                                 plan["elem_index"].data_ptr(), plan["n_elems"], Ti,
                                                 cabi.f64_array(dict_len), K, float(divisor), 1 if integral else 0,
                                                 div_flag, st), "fmlp_fedavg_multi_i64")
    return out


_pinned_pool: dict = {}


def _pinned_entry(lay, slot):
    """Pinned flat staging buffers of one client slot plus the per-tensor destination pointers / byte counts of the
    layout, pooled per (layout, slot): cudaHostAlloc of 28 MB costs milliseconds, and FedAvg's CPU path ends with a
    device-to-host read that synchronises, so a slot is free again when the next call starts."""
    key = (lay, slot)
    ent = _pinned_pool.get(key)
    if ent is None:
        pin_f = torch.zeros(max(lay.n_f32, 1), dtype=torch.float32, pin_memory=True)
        pin_i = torch.zeros(max(lay.n_i64, 1), dtype=torch.int64, pin_memory=True)
        is_int = np.array(lay.is_int, dtype=bool)
        offs = np.array(lay.offsets, dtype=np.int64)
        numels = np.array(lay.numels, dtype=np.int64)
        dsts = np.where(is_int, pin_i.data_ptr() + 8 * offs, pin_f.data_ptr() + 4 * offs).astype(np.int64)
        nbytes = np.where(is_int, 8 * numels, 4 * numels).astype(np.int64)
        ent = _pinned_pool[key] = (pin_f, pin_i, np.ascontiguousarray(dsts), np.ascontiguousarray(nbytes))
    return ent


def _stage_cpu_clients(w, dev):
    """Pack K CPU state_dicts into pinned flat buffers and upload them.  Per client: one pass for the 727 source
    pointers, ONE call of fmlp_host_copy_many (host threads, memory bandwidth) into the pooled pinned buffer, one
    asynchronous upload — the packing of client i+1 overlaps the upload of client i.  (Packing with per-tensor torch
    copies cost ~7 us of dispatch per tensor: 62 ms for 8 DenseNet121 clients, no faster than the reference's CPU
    FedAvg.)"""
    global _last_layout
    import os
    lib = cabi.lib()
    n_threads = max(1, min(16, (os.cpu_count() or 1)))
    staged = []
    for i, sd in enumerate(w):
        lay = _last_layout = layout_of(sd, like=_last_layout)
        pin_f, pin_i, dsts, nbytes = _pinned_entry(lay, i)
        vals = list(sd.values())
        if not all(map(torch.Tensor.is_contiguous, vals)):
            raise ValueError("FedAvg: non-contiguous state_dict tensor")
        if any(v.is_cuda for v in vals):
            raise TypeError("FedAvg: a state_dict mixes CPU and CUDA tensors")
        srcs = np.fromiter(map(torch.Tensor.data_ptr, vals), dtype=np.int64, count=len(vals))
        cabi.check(lib.fmlp_host_copy_many(srcs.ctypes.data, dsts.ctypes.data, nbytes.ctypes.data, len(vals), n_threads),
                   "fmlp_host_copy_many")
        d = FlatStateDict.empty(lay, dev, lazy=True, zero=False)      # only the flat buffers are used
        if lay.n_f32:
            d.flat_f32.copy_(pin_f[:lay.n_f32], non_blocking=True)
        if d.flat_i64 is not None:
            d.flat_i64.copy_(pin_i[:lay.n_i64], non_blocking=True)
        staged.append(d)
    return staged


def FedAvg(w, dict_len, _divide=True):
    """Weighted average of K client state_dicts (reference utils/FedAvg.py:7-14).

    w: list of K OrderedDict[str, Tensor] with identical keys; dict_len: K ints or floats.
    CUDA inputs are reduced in place on their device.  CPU inputs (the reference's stage 2 hands
    over `net.cpu()` weights, local_training.py:1251) are staged through pinned memory, reduced
    on the current CUDA device and returned as CPU tensors."""
    if len(w) == 0:
        raise ValueError("FedAvg of zero clients")
    if len(w) != len(dict_len):
        raise ValueError("FedAvg: len(w) != len(dict_len)")
    _check_same_keys(w)
    first = next(iter(w[0].values()))
    if first.is_cuda:
        for sd in w:
            cabi.require_cuda(*sd.values())
        return _fedavg_cuda(w, dict_len, _divide)
    if not torch.cuda.is_available():
        raise cabi.FedMLPNativeError("FedAvg needs a CUDA device (fedmlp_b200 has no CPU fallback)")
    dev = torch.device("cuda", torch.cuda.current_device())
    staged = _stage_cpu_clients(w, dev)
    res = _fedavg_cuda(staged, dict_len, _divide)
    host_flat = res.flat_f32.cpu()
    out = OrderedDict()
    lay = res.layout
    for i, k in enumerate(lay.keys):
        n = lay.numels[i]
        off = lay.offsets[i] if not lay.is_int[i] else lay.n_f32 + lay.offsets[i]
        out[k] = host_flat[off:off + n].view(lay.shapes[i])
    return out


def Fed_w(w, weight):
    """utils/FedAvg.py:16-23 — same body as FedAvg."""
    return FedAvg(w, weight)


def FedAvg_proto(Prototypes, weight, class_active_client_list, _rows_per_class=2):
    """Per-class weighted mean of the clients' prototypes (reference utils/FedAvg.py:72-93).
    Prototypes: list of K [2C, D] tensors; returns [2C, D] on the device of the inputs
    (the reference works on CPU tensors and returns a CPU tensor)."""
    K = len(Prototypes)
    if K == 0:
        raise ValueError("FedAvg_proto of zero clients")
    if K > cabi.MAX_CLIENTS:
        return _proto_avg_grouped(Prototypes, weight, class_active_client_list, _rows_per_class)
    p0 = Prototypes[0]
    was_cpu = not p0.is_cuda
    if was_cpu and not torch.cuda.is_available():
        raise cabi.FedMLPNativeError("FedAvg_proto needs a CUDA device (no CPU fallback)")
    dev = p0.device if p0.is_cuda else torch.device("cuda", torch.cuda.current_device())
    stacked = torch.stack([p.to(dev, dtype=torch.float32) for p in Prototypes]).contiguous()
    rpc = _rows_per_class
    C2, D = stacked.shape[1], stacked.shape[2]
    C = C2 // rpc
    if len(class_active_client_list) > C:
        raise ValueError("class_active_client_list longer than the number of classes")
    masks = [0] * C
    for cls, clients in enumerate(class_active_client_list):
        for cid in clients:
            if not 0 <= int(cid) < K:
                raise IndexError(f"client id {cid} out of range")
            masks[cls] |= 1 << int(cid)
    out = torch.zeros(C2, D, dtype=torch.float32, device=dev)
    # the reference only fills the rows of the classes it iterates over; the others stay 0
    n_listed = len(class_active_client_list)
    with torch.cuda.device(dev):
        tmp = torch.empty(C2, D, dtype=torch.float32, device=dev)
        cabi.check(cabi.lib().fmlp_proto_avg_f32(stacked.data_ptr(), K, C, D, rpc, cabi.f64_array(weight),
                                                 cabi.u64_array(masks), tmp.data_ptr(), cabi.stream_ptr(dev)),
                   "fmlp_proto_avg_f32")
        out[:rpc * n_listed] = tmp[:rpc * n_listed]
    return out.cpu() if was_cpu else out


def _proto_avg_grouped(Prototypes, weight, class_active_client_list, rpc):
    """More than 64 clients: the kernel's client masks are 64 bits wide, so the clients are averaged in groups
    of 64 and the group means are combined per class with the groups' class weights by the same kernel.
    Same value as the reference up to fp32 summation association (<= 1e-6 relative), NaN for an empty class."""
    K = len(Prototypes)
    n_listed = len(class_active_client_list)
    G = cabi.MAX_CLIENTS
    groups = [range(g0, min(K, g0 + G)) for g0 in range(0, K, G)]
    means, wsum = [], np.zeros((len(groups), n_listed), dtype=np.float64)
    for gi, g in enumerate(groups):
        sub_lists = [[int(c) - g.start for c in clients if g.start <= int(c) < g.stop] for clients in class_active_client_list]
        for cls, clients in enumerate(class_active_client_list):
            wsum[gi, cls] = float(sum(weight[int(c)] for c in clients if g.start <= int(c) < g.stop))
        means.append(FedAvg_proto([Prototypes[i] for i in g], [weight[i] for i in g], sub_lists, rpc))
    out = torch.zeros_like(means[0])
    for cls in range(n_listed):
        members = [gi for gi in range(len(groups)) if wsum[gi, cls] > 0]
        rows = slice(rpc * cls, rpc * cls + rpc)
        if not members:
            out[rows] = float("nan")           # 0/0 in the reference (FedAvg.py:85-86)
            continue
        lvl2 = FedAvg_proto([means[gi][rows].contiguous() for gi in range(len(groups))], list(wsum[:, cls]), [members], rpc)
        out[rows] = lvl2[:rpc]
    return out


def FedAvg_rela(Prototypes, weight, class_active_client_list):
    """utils/FedAvg.py:95-103: like FedAvg_proto with ONE row per class ([C, D] inputs)."""
    return FedAvg_proto(Prototypes, weight, class_active_client_list, _rows_per_class=1)


def model_dist(w_1, w_2):
    """sum over the float tensors, in key order, of ||w_1[k] - w_2[k]||_2 as a Python float
    (reference utils/FedNoRo.py:106-115; utils/FedAvg.py:42-49 is the same loop without the int64
    skip, which makes torch.norm raise on BatchNorm counters — the skip is kept here).
    One deterministic multi-tensor launch pair instead of 2 x 727 tiny ops and 727 host syncs."""
    if w_1.keys() != w_2.keys():
        raise AssertionError("Error: cannot compute distance between dict with different keys")
    first = next(iter(w_1.values()))
    if not first.is_cuda:
        if not torch.cuda.is_available():
            raise cabi.FedMLPNativeError("model_dist needs a CUDA device (no CPU fallback)")
        dev = torch.device("cuda", torch.cuda.current_device())
        w_1 = OrderedDict((k, v.to(dev)) for k, v in w_1.items())
        w_2 = OrderedDict((k, v.to(dev)) for k, v in w_2.items())
    cabi.require_cuda(*w_1.values(), *w_2.values())
    layout = layout_of(w_1)
    dev = next(iter(w_1.values())).device
    plan = _multi_plan(layout, dev)
    f_idx = plan["f_idx"]           # the float tensors of w_1: the reference skips on w_1's dtype (FedNoRo.py:110-111)
    T = len(f_idx)
    v1, v2 = list(w_1.values()), list(w_2.values())
    # contiguous float32 operands that OUTLIVE the launch and the .item() sync below (a temporary made by
    # .contiguous() / .float() could be recycled by the allocator before the kernel runs)
    a_ops, b_ops = [], []
    for t in f_idx:
        a, b = v1[t], v2[t]
        if a.shape != b.shape:
            raise ValueError("model_dist: the two state_dicts have different shapes")
        # w_2 may be a FedAvg output, whose int64 counters became float32: only w_1's float keys matter
        a_ops.append(a.contiguous() if a.dtype == torch.float32 else a.float().contiguous())
        b_ops.append(b.contiguous() if b.dtype == torch.float32 else b.float().contiguous())
    table = np.array([[x.data_ptr() for x in a_ops], [x.data_ptr() for x in b_ops]], dtype=np.int64).reshape(-1)
    lib = cabi.lib()
    with torch.cuda.device(dev):
        table_dev = torch.from_numpy(table).to(dev)
        out = torch.empty(1, dtype=torch.float32, device=dev)
        ws = torch.empty(max(lib.fmlp_model_dist_ws_bytes(plan["n_chunks"], T), 256), dtype=torch.uint8, device=dev)
        cabi.check(lib.fmlp_model_dist_f32(table_dev.data_ptr(), table_dev.data_ptr() + 8 * T, plan["numel"].data_ptr(),
                                           plan["chunk_tensor"].data_ptr(), plan["chunk_start"].data_ptr(),
                                           plan["tensor_chunk0"].data_ptr(), plan["n_chunks"], T, out.data_ptr(),
                                           ws.data_ptr(), ws.numel(), cabi.stream_ptr(dev)), "fmlp_model_dist_f32")
        result = float(out.item())
    del a_ops, b_ops
    return result


def RSCFed(DMA, w_locals, K, dict_len, M):
    """utils/FedAvg.py:25-40: per sub-sampled group, a plain average, distance-aware re-weighting
    (exp(-0.01 * dist / n)) and a weighted average; then the plain average of the M group models."""
    import math

    w_sub = []
    for group in DMA:
        w_select = [w_locals[i] for i in group]
        n_total = sum(dict_len[i] for i in group)
        w_avg = Fed_w(w_select, [1] * K)
        w = []
        for i in group:
            a = dict_len[i] / n_total
            b = math.exp((-0.01) * (model_dist(w_locals[i], w_avg) / dict_len[i]))
            w.append(a * b)
        w_sub.append(Fed_w(w_select, w))
    return Fed_w(w_sub, [1] * M)


def DaAgg(w, dict_len, clean_clients, noisy_clients):
    """utils/FedNoRo.py:84-103: distance-aware aggregation.  Noisy clients are down-weighted by
    exp(-d / d_max), d = distance to the nearest clean client; the result is the weighted SUM with
    the renormalised weights (no further division)."""
    client_weight = np.array(dict_len)
    client_weight = client_weight / client_weight.sum()
    distance = np.zeros(len(dict_len))
    for n_idx in noisy_clients:
        distance[n_idx] = min(model_dist(w[n_idx], w[c_idx]) for c_idx in clean_clients)
    distance = distance / distance.max()
    client_weight = client_weight * np.exp(-distance)
    client_weight = client_weight / client_weight.sum()
    return FedAvg(w, [float(x) for x in client_weight], _divide=False)


def FedAvg_tao(t, weight, class_active_client_list=None):
    """Per-class weighted mean of the clients' difficulty statistics t (reference
    utils/FedAvg.py:51-70).  t: K float64 arrays [C]; returns a float64 numpy [C] like the
    reference.  K*C doubles go through fmlp_tao_avg_f64 (IEEE double, the reference's operation
    order -> bit-identical); main.py:223 passes the per-class lists of clients that MISS the class."""
    K, n = len(t), len(t[0])
    if K > cabi.MAX_CLIENTS:
        return _tao_avg_grouped(t, weight, class_active_client_list)
    if not torch.cuda.is_available():
        raise cabi.FedMLPNativeError("FedAvg_tao needs a CUDA device (no CPU fallback)")
    dev = torch.device("cuda", torch.cuda.current_device())
    t_dev = torch.from_numpy(np.ascontiguousarray(np.stack([np.asarray(x, dtype=np.float64) for x in t]))).to(dev)
    out = torch.empty(n, dtype=torch.float64, device=dev)
    masks = None
    if class_active_client_list is not None:
        m = [0] * n
        for cls, clients in enumerate(class_active_client_list):
            for cid in clients:
                if 0 <= int(cid) < K:
                    m[cls] |= 1 << int(cid)
        masks = cabi.u64_array(m)
    with torch.cuda.device(dev):
        cabi.check(cabi.lib().fmlp_tao_avg_f64(t_dev.data_ptr(), K, n, cabi.f64_array([float(x) for x in weight[:K]]), masks,
                                               float(sum(weight)), out.data_ptr(), cabi.stream_ptr(dev)), "fmlp_tao_avg_f64")
    return out.cpu().numpy()


def _tao_avg_grouped(t, weight, class_client_list):
    """More than 64 clients: per-group results of the kernel combined per class by the same kernel (float64;
    equal to the reference up to summation association, ~1e-16 relative)."""
    K, n = len(t), len(t[0])
    G = cabi.MAX_CLIENTS
    groups = [range(g0, min(K, g0 + G)) for g0 in range(0, K, G)]
    lists = class_client_list if class_client_list is not None else [list(range(K))] * n
    part = np.zeros((len(groups), n), dtype=np.float64)
    wsum = np.zeros((len(groups), n), dtype=np.float64)
    for gi, g in enumerate(groups):
        sub = [[int(c) - g.start for c in clients if g.start <= int(c) < g.stop] for clients in lists]
        for cls, clients in enumerate(lists):
            wsum[gi, cls] = float(sum(float(weight[int(c)]) for c in clients if g.start <= int(c) < g.stop))
        part[gi] = FedAvg_tao([t[i] for i in g], [weight[i] for i in g], sub)
    out = np.ones(n, dtype=np.float64)
    for cls in range(n):
        members = [gi for gi in range(len(groups)) if wsum[gi, cls] > 0]
        if members:
            out[cls] = FedAvg_tao([part[gi, cls:cls + 1] for gi in range(len(groups))], list(wsum[:, cls]), [members])[0]
    return out      # without lists utils/FedAvg.py:56-59 divides by the weight of ALL clients: the grouped mean does too
