"""FedAvg aggregation sweep (BASELINE.json configs[4]): K clients x P fp32 parameters on one GPU.
Prints one JSON line per point: algorithmic GB/s = (K+1)*4*P / t (CUDA events, median of 7)."""
import json, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import fedmlp_b200 as F

dev = torch.device("cuda", 0)
peak = 6447.8
try:
    peak = float(json.load(open(Path(__file__).resolve().parent.parent / "MEASURED_PEAKS.json"))["hbm_gbs"])
except Exception:
    pass
budget = 120e9
rows = []
for P in (8_000_000, 25_000_000, 50_000_000, 100_000_000):
    for K in (8, 16, 32, 64, 128, 256, 512):
        if K * P * 4 > budget:
            continue
        bufs = [torch.empty(P, dtype=torch.float32, device=dev) for _ in range(K)]
        for i, b in enumerate(bufs):
            b.normal_(0, 0.02)
        w = [5000 + i for i in range(K)]
        out = torch.empty(P, dtype=torch.float32, device=dev)
        for _ in range(2):
            F.fedavg_flat_buffers(bufs, w, out=out)
        ts = []
        for _ in range(7):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); F.fedavg_flat_buffers(bufs, w, out=out); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort(); ms = ts[len(ts) // 2]
        alg = (K + 1) * 4 * P
        launches = (K + 63) // 64
        moved = alg + 2 * (launches - 1) * 4 * P      # chained launches re-read and re-write the accumulator
        row = dict(K=K, P=P, ms=round(ms, 4), alg_gbs=round(alg / ms / 1e6, 1), moved_gbs=round(moved / ms / 1e6, 1),
                   frac_of_peak=round(alg / ms / 1e6 / peak, 3), launches=launches)
        rows.append(row); print(json.dumps(row), flush=True)
        del bufs, out
        torch.cuda.empty_cache()
