"""FlatAdam (one launch) vs torch.optim.Adam on a DenseNet121-shaped parameter set (CUDA events)."""
import sys, json
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from torch import nn
from fedmlp_b200.shapes import densenet121_state_shapes
from fedmlp_b200.optim import FlatAdam

class Bag(nn.Module):
    def __init__(self, shapes):
        super().__init__()
        for i, (k, (shp, dt)) in enumerate(shapes.items()):
            name = k.replace(".", "_")
            if dt == torch.int64: self.register_buffer(name, torch.zeros(shp, dtype=dt))
            elif "running" in k: self.register_buffer(name, torch.zeros(shp))
            else: setattr(self, name, nn.Parameter(torch.randn(shp) * 0.02))

def timeit(fn, iters=30):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

shapes = densenet121_state_shapes(5)
res = {}
for kind in ("flat", "torch_foreach", "torch_fused"):
    torch.manual_seed(0)
    net = Bag(shapes).cuda()
    n_param = sum(p.numel() for p in net.parameters())
    if kind == "flat":
        opt = FlatAdam(net, lr=3e-5, weight_decay=5e-4)
        opt.grad.normal_()
        fn = lambda: opt.step()
    else:
        for p in net.parameters(): p.grad = torch.randn_like(p)
        opt = torch.optim.Adam(net.parameters(), lr=3e-5, weight_decay=5e-4, foreach=(kind == "torch_foreach"), fused=(kind == "torch_fused"))
        fn = lambda: opt.step()
    ms = timeit(fn)
    res[kind] = {"ms": round(ms, 4), "gbs": round(28 * n_param / ms / 1e6, 1)}
res["params"] = n_param
print(json.dumps(res))
