"""Fused fold + all-reduce kernel with a single rank (no peers) vs the single-GPU FedAvg kernel: isolates the
fold phase of fedavg_allreduce.cu (also the ncu target for it)."""
import json, os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import torch.distributed as dist
import fedmlp_b200 as F
from fedmlp_b200 import dist as fd

os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29544")
dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
dist.init_process_group("nccl", rank=0, world_size=1, device_id=dev)
def timed(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for K, P in ((8, 7042752), (64, 25_000_000)):
    bufs = [torch.empty(P, dtype=torch.float32, device=dev).normal_(0, 0.02) for _ in range(K)]
    w = [5000 + i for i in range(K)]; wn = [x / sum(w) for x in w]
    out = torch.empty(P, dtype=torch.float32, device=dev)
    row = {"K": K, "P": P, "flat_ms": round(timed(lambda: F.fedavg_flat_buffers(bufs, w, out=out)), 4)}
    for nc in (1, 2):
        fused = fd.FusedFedAvgAllReduce(P, device=dev, n_chunks=nc)
        row[f"fused_nc{nc}_ms"] = round(timed(lambda: fused(bufs, wn)), 4)
        ref = F.fedavg_flat_buffers(bufs, w, out=out)
        row[f"rel_err_nc{nc}"] = float((fused(bufs, wn) - ref).abs().max() / ref.abs().max())
    row["alg_gb"] = (K + 1) * 4 * P / 1e9
    print(json.dumps(row), flush=True)
dist.destroy_process_group()
