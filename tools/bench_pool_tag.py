"""Fused pooling + tagging (csrc/pool_tag.cu) vs the unfused chain (torch relu + adaptive_avg_pool2d,
then fmlp_tag_sim_f32 on the materialised features).  CUDA events; buffers are rotated so that the
bytes touched between two uses of a buffer exceed L2 (126 MB)."""
import sys, json
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import torch.nn.functional as F
from fedmlp_b200 import pooling, tagging

def timeit(fn, iters):
    for i in range(3): fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

out = []
cases = [(32, 1024, 5, "nchw"), (128, 1024, 5, "nchw"), (128, 1024, 5, "nhwc"), (128, 1280, 14, "nchw"),
         (128, 1280, 14, "nhwc"), (2048, 1024, 5, "nchw"), (2048, 1024, 5, "nhwc"), (2048, 1280, 14, "nchw")]
if len(sys.argv) > 1:
    cases = [c for c in cases if str(c[0]) in sys.argv[1:]]
for B, D, C, layout in cases:
    bytes_per = B * D * 49 * 4
    nbuf = max(2, min(16, int(400e6 // bytes_per) + 1))
    torch.manual_seed(0)
    bufs = [torch.randn(B, D, 7, 7, device="cuda") for _ in range(nbuf)]
    if layout == "nhwc":
        bufs = [b.contiguous(memory_format=torch.channels_last) for b in bufs]
    proto = torch.rand(2 * C, D, device="cuda") + 0.05
    missing = list(range(1, C))
    table = pooling.build_sim_table(proto, missing, "folded")
    sim = torch.empty(C, B, device="cuda")
    feat = torch.empty(B, D, device="cuda")
    seg = [0, B]
    def fused(i):
        pooling.pool_tag(bufs[i % nbuf], table, sim_out=sim, feat_out=feat)
    def pool_only(i):
        pooling.pool_tag(bufs[i % nbuf], feat_out=feat)
    def unfused(i):
        f = F.adaptive_avg_pool2d(F.relu(bufs[i % nbuf]), (1, 1)).flatten(1)
        tagging.tag_similarity(f, proto, [missing], seg, out=sim, mode="folded")
    iters = 200 if B <= 128 else 20
    r = {"B": B, "D": D, "C": C, "layout": layout, "MB": round(bytes_per / 1e6, 1)}
    for name, fn in (("fused", fused), ("pool_only", pool_only), ("torch_plus_sim", unfused)):
        ms = timeit(fn, iters)
        r[name] = {"us": round(ms * 1e3, 1), "gbs": round((bytes_per + B * D * 4) / ms / 1e6, 1)}
    print(json.dumps(r), flush=True)
