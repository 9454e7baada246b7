"""The work-queue aggregation kernel with a single rank (no peers): isolates the queue / fold / local reduce cost
against the single-GPU FedAvg kernel (also the ncu target for it)."""
import json, os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import torch.distributed as dist
import fedmlp_b200 as F
from fedmlp_b200 import dist as fd

os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29545")
dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
dist.init_process_group("nccl", rank=0, world_size=1, device_id=dev)


def timed(fn, n=30):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


K, P = 8, 7042752
bufs = [torch.empty(P, dtype=torch.float32, device=dev).normal_(0, 0.02) for _ in range(K)]
w = [5000 + i for i in range(K)]; wn = [x / sum(w) for x in w]
out = torch.empty(P, dtype=torch.float32, device=dev)
print(json.dumps({"kind": "fedavg_flat_kernel", "ms": round(timed(lambda: F.fedavg_flat_buffers(bufs, w, out=out)), 4)}), flush=True)
only = sys.argv[1:] and sys.argv[1] == "one"
for nc, fi, ri, ctas in ([(4, 2, 2, 0)] if only else [(1, 2, 2, 0), (4, 2, 2, 0), (4, 4, 4, 0), (4, 8, 8, 0), (4, 16, 16, 0), (2, 32, 32, 0), (8, 2, 2, 0), (4, 2, 2, 74)]):
    q = fd.QueuedAggregation(P, 0, 0, device=dev, n_chunks=nc, fold_iters=fi, red_iters=ri, max_ctas=ctas)
    ms = timed(lambda: q(bufs, wn))
    ref = F.fedavg_flat_buffers(bufs, w, out=out)
    err = float((q(bufs, wn)[0] - ref).abs().max() / ref.abs().max())
    print(json.dumps({"kind": "queue world=1", "chunks": nc, "fold_iters": fi, "red_iters": ri, "max_ctas": ctas, "ms": round(ms, 4), "rel_err": err}), flush=True)
    del q
dist.destroy_process_group()
