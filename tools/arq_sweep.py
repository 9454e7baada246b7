"""Sweep of the work-queue aggregation kernel (csrc/fedavg_allreduce_q.cu) on N GPUs (torchrun): chunks x fold / reduce
work-item sizes x NVLS / peer-to-peer, against the round-1 cooperative kernel and fold + NCCL.  One JSON line per
setting: max-over-ranks CUDA-event time (median of 7, barrier before each), correctness vs the NCCL path."""
import json, os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import torch.distributed as dist
from fedmlp_b200 import dist as fd

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
P, K, C, D, J = 7042752, 8, 5, 1024, 121
if len(sys.argv) > 1:
    P, K = int(sys.argv[1]), int(sys.argv[2])
T, M = 2 * C * D, 3 * C + J
bufs = [torch.empty(P, dtype=torch.float32, device=dev).normal_(0, 0.02) for _ in range(K)]
tails = [torch.randn(T, device=dev) for _ in range(K)]
tail64 = torch.rand(M, dtype=torch.float64, device=dev)
weights = [5000 + rank * K + i for i in range(K)]
tot = torch.tensor([float(sum(weights))], dtype=torch.float64, device=dev)
dist.all_reduce(tot)
wn = [w / float(tot.item()) for w in weights]


def timed(fn, n=7):
    ts = []
    for _ in range(n):
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    t = torch.tensor([ts[len(ts) // 2]], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


out = torch.empty(P, dtype=torch.float32, device=dev)
ref = fd.fedavg_flat_distributed(bufs, weights, total_weight=float(tot.item()), out=out).clone()
ms_nccl = timed(lambda: fd.fedavg_flat_distributed(bufs, weights, total_weight=float(tot.item()), out=out))
if rank == 0:
    print(json.dumps(dict(n_gpus=world, kind="fold + nccl all_reduce", ms=round(ms_nccl, 4))), flush=True)
try:
    r01 = fd.FusedFedAvgAllReduce(P, device=dev)
    for _ in range(2):
        r01(bufs, wn)
    ms = timed(lambda: r01(bufs, wn))
    if rank == 0:
        print(json.dumps(dict(n_gpus=world, kind="round-1 cooperative kernel", ms=round(ms, 4))), flush=True)
    del r01
except Exception as exc:
    if rank == 0:
        print(json.dumps(dict(kind="round-1 kernel", error=str(exc)[:200])), flush=True)
torch.cuda.empty_cache()
settings = [(mc, nc, fi, ri, ctas) for mc in (1, 0) for nc in (2, 4, 8) for (fi, ri) in ((2, 2), (2, 1), (4, 2), (4, 4)) for ctas in (0,)]
settings += [(1, 4, 2, 2, 96), (1, 4, 2, 2, 64), (1, 1, 2, 2, 0), (1, 16, 2, 2, 0)]
if os.environ.get("FMLP_SWEEP_SHORT"):
    settings = [(mc, nc, fi, ri, 0) for mc in (1, 0) for nc in (2, 4, 8) for (fi, ri) in ((2, 2), (4, 4))] + [(1, 1, 2, 2, 0), (1, 4, 8, 8, 0)]
for mc, nc, fi, ri, ctas in settings:
    try:
        q = fd.QueuedAggregation(P, T, M, device=dev, n_chunks=nc, use_multicast=bool(mc), fold_iters=fi, red_iters=ri, max_ctas=ctas)
        for _ in range(2):
            res, rt, r64 = q(bufs, wn, tail_bufs=tails, tail_f64=tail64)
        torch.cuda.synchronize()
        err = float((res - ref).abs().max() / ref.abs().max())
        ms = timed(lambda: q(bufs, wn, tail_bufs=tails, tail_f64=tail64))
        if rank == 0:
            print(json.dumps(dict(n_gpus=world, kind="queue", path=q.path, chunks=nc, fold_iters=fi, red_iters=ri, max_ctas=ctas,
                                  ms=round(ms, 4), rel_err_vs_nccl=err)), flush=True)
        del q, res, rt, r64
        torch.cuda.empty_cache()
    except Exception as exc:
        if rank == 0:
            print(json.dumps(dict(kind="queue", mc=mc, chunks=nc, error=f"{type(exc).__name__}: {exc}"[:200])), flush=True)
dist.barrier()
dist.destroy_process_group()
