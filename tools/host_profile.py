"""Host-side profile of the round hot path (where does Python/launch time go?)."""
import cProfile, pstats, sys, time, io
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import bench
from fedmlp_b200.round import ClientShard
from fedmlp_b200 import _cabi as cabi

a = bench.parse_args()
dev = torch.device("cuda", 0)
w = bench.Workload(a.config, a)
inp = bench.make_device_inputs(w, 0, dev)
S, C = w.S, w.C
shard = ClientShard([w.n] * S, C, [[k % C] for k in range(S)], device=dev)
fed_out = torch.empty(inp["Ppad"], dtype=torch.float32, device=dev)
def step():
    return shard.round_hot_path(inp["feat_tag"], inp["proto"], inp["logits"], inp["logits_glob"], inp["labels"],
                                inp["feat_proto"], inp["logits_proto"], inp["flats"], inp["weights"], fedavg_out=fed_out)
for _ in range(5): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20): step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host enqueue per step {(t1-t0)/20*1e3:.3f} ms; incl. drain {(t2-t0)/20*1e3:.3f} ms")
# time individual C calls
lib = cabi.lib()
names = ["fmlp_tag_sim_f32", "fmlp_tag_select", "fmlp_mask_fill", "fmlp_loss_stage2_f32", "fmlp_proto_build_f32", "fmlp_fedavg_flat_f32"]
acc = {n: [0.0, 0] for n in names}
class Wrap:
    def __init__(self, fn, n): self.fn, self.n = fn, n
    def __call__(self, *args):
        t = time.perf_counter(); r = self.fn(*args); acc[self.n][0] += time.perf_counter() - t; acc[self.n][1] += 1; return r
class LibProxy:
    def __getattr__(self, n):
        f = getattr(lib, n)
        return Wrap(f, n) if n in acc else f
cabi._lib = LibProxy()
for _ in range(20): step()
torch.cuda.synchronize()
for n, (t, c) in acc.items():
    print(f"{n:28s} calls={c:4d}  host {t/c*1e6:9.1f} us/call")
cabi._lib = lib
pr = cProfile.Profile(); pr.enable()
for _ in range(20): step()
pr.disable(); torch.cuda.synchronize()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(25); print(s.getvalue()[:5000])
