"""Multi-GPU FedAvg aggregation sweep (BASELINE.json configs[4]): K clients x P fp32 parameters sharded over the
ranks of one box (torchrun).  Per point, the fused fold + all-reduce kernel and the fold + NCCL all-reduce path:
whole-job algorithmic GB/s = (K+1)*4*P / t, t = max over ranks (CUDA events, median of 5, barrier before each)."""
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
points = [(K, P) for P in (8_000_000, 25_000_000, 100_000_000) for K in (64, 256, 512)]
if len(sys.argv) > 1:
    points = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]]


def timed(fn):
    ts = []
    for _ in range(5):
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    t = torch.tensor([ts[len(ts) // 2]], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


fused_by_p = {}
for K, P in points:
    k_local = K // world
    if k_local < 1 or k_local > 64:
        continue
    bufs = [torch.empty(P, dtype=torch.float32, device=dev).normal_(0, 0.02) for _ in range(k_local)]
    weights = [5000 + rank * k_local + i for i in range(k_local)]
    tot = torch.tensor([float(sum(weights))], dtype=torch.float64, device=dev)
    dist.all_reduce(tot)
    wn = [w / float(tot.item()) for w in weights]
    if P not in fused_by_p:
        fused_by_p.clear()                       # one set of symmetric buffers at a time
        fused_by_p[P] = fd.FusedFedAvgAllReduce(P, device=dev)
    fused = fused_by_p[P]
    out = torch.empty(P, dtype=torch.float32, device=dev)
    for _ in range(2):
        res_f = fused(bufs, wn)
        res_n = fd.fedavg_flat_distributed(bufs, weights, total_weight=float(tot.item()), out=out)
    torch.cuda.synchronize()
    err = float((res_f - res_n).abs().max() / res_n.abs().max())
    chk = torch.tensor([float(res_f.double().sum())], dtype=torch.float64, device=dev)
    lo, hi = chk.clone(), chk.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    ms_f = timed(lambda: fused(bufs, wn))
    ms_n = timed(lambda: fd.fedavg_flat_distributed(bufs, weights, total_weight=float(tot.item()), out=out))
    alg = (K + 1) * 4 * P
    if rank == 0:
        print(json.dumps(dict(n_gpus=world, K=K, P=P, clients_per_gpu=k_local, fused_ms=round(ms_f, 4), nccl_path_ms=round(ms_n, 4),
                              fused_alg_gbs=round(alg / ms_f / 1e6, 1), nccl_path_alg_gbs=round(alg / ms_n / 1e6, 1),
                              fused_per_gpu_gbs=round(alg / ms_f / 1e6 / world, 1), fused_vs_nccl_rel_err=err,
                              identical_on_all_ranks=bool(float(lo.item()) == float(hi.item())))), flush=True)
    del bufs, out
    torch.cuda.empty_cache()
dist.destroy_process_group()
