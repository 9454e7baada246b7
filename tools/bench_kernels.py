"""Micro-benchmark of the individual kernels with CUDA events (tuning aid, not the bench)."""
import argparse, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
ap = argparse.ArgumentParser()
ap.add_argument("--lib", default=None)
ap.add_argument("--iters", type=int, default=50)
ap.add_argument("--classes", type=int, default=5)
ap.add_argument("--dim", type=int, default=1024)
ap.add_argument("--rows", type=int, default=6875)
ap.add_argument("--clients", type=int, default=8)
ap.add_argument("--mode", default="pair")
args = ap.parse_args()
import torch
from fedmlp_b200 import _cabi
if args.lib:
    _cabi._lib = _cabi.load(Path(args.lib))
import bench
import fedmlp_b200 as F
from fedmlp_b200.round import ClientShard
class A: pass
a = A(); a.clients_per_gpu = args.clients; a.rows_per_client = args.rows; a.classes = args.classes; a.dim = args.dim; a.sim_mode = args.mode
dev = torch.device("cuda", 0)
inp = bench.make_device_inputs(a, 0, dev)
S, C = a.clients_per_gpu, a.classes
shard = ClientShard([a.rows_per_client] * S, C, [[k % C] for k in range(S)], device=dev, sim_mode=a.sim_mode)
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
def timeit(fn, name, nbytes):
    for _ in range(3): fn()
    ts = []
    for _ in range(args.iters):
        flush.add_(1.0)  # L2 flush between iterations
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort(); med = ts[len(ts)//2]
    print(f"{name:14s} median {med*1e3:8.1f} us  min {ts[0]*1e3:8.1f} us  -> {nbytes/med/1e6:8.1f} GB/s (median)")
ab = bench.alg_bytes(a, inp)
timeit(lambda: shard.tagger.similarity(inp["feat_tag"], inp["proto"], a.sim_mode), "sim", ab["sim"])
timeit(lambda: shard.tagger.select(0.005, 0.01), "select", ab["select_fill"])
timeit(lambda: shard.tagger.fill(inp["labels"]), "fill", 3*4*inp["N"]*C)
timeit(lambda: F.build_prototypes(inp["feat_proto"], inp["labels"], inp["logits_proto"], shard.active, shard.missing, 0.3, 0.7, True, seg_rows=shard.seg_rows), "proto", ab["proto"])
fed_out = torch.empty(inp["Ppad"], dtype=torch.float32, device=dev)
timeit(lambda: F.fedavg_flat_buffers(inp["flats"], inp["weights"], out=fed_out), "fedavg", ab["fedavg"])
y, distill, sup = shard.tagger.fill(inp["labels"])
from fedmlp_b200.losses import launch_stage2
loss = torch.empty(1, device=dev); dz = torch.empty_like(inp["logits"])
n = a.rows_per_client
timeit(lambda: launch_stage2(inp["logits"][:n], inp["logits_glob"][:n], y[:n], distill[:n], 0, loss, dz[:n]), "loss(1 client)", 5*4*n*C)
timeit(lambda: launch_stage2(inp["logits"], inp["logits_glob"], y, distill, 0, loss, dz), "loss(all rows)", 5*4*inp["N"]*C)
