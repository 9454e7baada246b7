"""Per-item timeline of the work-queue aggregation kernel (debug build tools/bin/lib_arq_trace.so, FEDMLP_B200_LIB):
torchrun on N GPUs; rank 0 prints, per item kind and chunk, when the items were published / finished / signalled."""
import ctypes, json, os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch
import torch.distributed as dist
from fedmlp_b200 import _cabi as cabi, dist as fd

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", local); torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
lib = cabi.load()
dump = ctypes.CDLL(str(cabi.LIB_PATH)).fmlp_arq_trace_dump
dump.restype = ctypes.c_int; dump.argtypes = [ctypes.c_void_p, ctypes.c_int]
P, K = 7042752, 8
bufs = [torch.empty(P, dtype=torch.float32, device=dev).normal_(0, 0.02) for _ in range(K)]
wn = [1.0 / (K * world)] * K
for mc in (1, 0):
    for nc in (4, 1):
        q = fd.QueuedAggregation(P, 0, 0, device=dev, n_chunks=nc, use_multicast=bool(mc))
        buf = np.zeros((1 << 16, 4), dtype=np.uint64)
        for it in range(4):
            torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
            dump(buf.ctypes.data, 1 << 16)          # reset
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); q(bufs, wn); e1.record(); torch.cuda.synchronize()
        n = dump(buf.ctypes.data, 1 << 16)
        rec = buf[:n]
        kind = (rec[:, 0] >> np.uint64(32)) & np.uint64(0xff)
        chunk = (rec[:, 0] >> np.uint64(40)) & np.uint64(0xff)
        t0 = rec[:, 1].min()
        rows = []
        for kd, nm in ((0, "F"), (1, "R"), (2, "FINAL")):
            for c in range(nc):
                m = (kind == kd) & (chunk == c)
                if not m.any():
                    continue
                pub, done, sig = rec[m, 1] - t0, rec[m, 2] - t0, rec[m, 3] - t0
                rows.append(f"{nm}{c}: n={int(m.sum()):4d} pub {pub.min()/1e3:7.1f}..{pub.max()/1e3:7.1f}us  done {done.min()/1e3:7.1f}..{done.max()/1e3:7.1f}us  "
                            f"dur avg {float((done - pub).mean())/1e3:6.2f} max {float((done - pub).max())/1e3:6.2f}us  sig lag avg {float((sig - done).mean())/1e3:5.2f}us")
        if rank == 0:
            print(f"== world {world} path {q.path} chunks {nc}: kernel {e0.elapsed_time(e1)*1e3:.1f} us, {n} records", flush=True)
            print("\n".join(rows), flush=True)
        del q
        torch.cuda.empty_cache()
dist.barrier(); dist.destroy_process_group()
