"""Run each hot-path kernel a few times on bench-sized inputs (target of `ncu -k regex:...`)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import bench
from fedmlp_b200.round import ClientShard

a = bench.parse_args()
dev = torch.device("cuda", 0)
inp = bench.make_device_inputs(a, 0, dev)
S, C = a.clients_per_gpu, a.classes
shard = ClientShard([a.rows_per_client] * S, C, [[k % C] for k in range(S)], device=dev, sim_mode=a.sim_mode)
fed_out = torch.empty(inp["Ppad"], dtype=torch.float32, device=dev)
for _ in range(3):
    shard.round_hot_path(inp["feat_tag"], inp["proto"], inp["logits"], inp["logits_glob"], inp["labels"],
                         inp["feat_proto"], inp["logits_proto"], inp["flats"], inp["weights"], fedavg_out=fed_out)
torch.cuda.synchronize()
print("done")
