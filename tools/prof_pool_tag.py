"""Target for `ncu -k regex:pool_tag`: three launches at B=2048 — pool only, fused C=5, fused C=14 (D=1280)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from fedmlp_b200 import pooling
torch.manual_seed(0)
B = 2048
for D, C, score in ((1024, 5, False), (1024, 5, True), (1280, 14, True)):
    x = torch.randn(B, D, 7, 7, device="cuda")
    proto = torch.rand(2 * C, D, device="cuda") + 0.05
    table = pooling.build_sim_table(proto, list(range(1, C)), "folded") if score else None
    sim = torch.empty(C, B, device="cuda")
    feat = torch.empty(B, D, device="cuda")
    pooling.pool_tag(x, table, sim_out=sim, feat_out=feat)
    torch.cuda.synchronize()
    del x
print("done")
