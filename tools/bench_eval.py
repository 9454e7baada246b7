"""On-device evaluation metrics vs the sklearn/numpy path of the reference (timed on the host)."""
import sys, json, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from fedmlp_b200 import evaluation
from oracle import fedmlp_oracle as O   # checker / CPU baseline only

N, C = 25596, 14     # ChestX-ray14 official test split
g = torch.Generator().manual_seed(0)
labels = (torch.rand(N, C, generator=g) < torch.linspace(0.02, 0.2, C)).float()
logits = 1.5 * torch.randn(N, C, generator=g) + 1.2 * (labels - 0.5)
dl, dy = logits.cuda(), labels.cuda()
for _ in range(3): evaluation.class_statistics(dl, dy)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): evaluation.class_statistics(dl, dy)
e1.record(); torch.cuda.synchronize()
gpu_ms = e0.elapsed_time(e1) / 20
t0 = time.perf_counter(); r = evaluation.multilabel_metrics(dl, dy); torch.cuda.synchronize(); api_ms = (time.perf_counter() - t0) * 1e3
probs = torch.sigmoid(dl).cpu().numpy()
t0 = time.perf_counter(); ref = O.eval_metrics(probs, labels.numpy()); cpu_ms = (time.perf_counter() - t0) * 1e3
print(json.dumps({"N": N, "C": C, "device_kernels_ms": round(gpu_ms, 4), "api_call_ms": round(api_ms, 3), "cpu_port_ms": round(cpu_ms, 1),
                  "max_abs_diff": max(abs(float(r[k]) - float(ref[k])) for k in ("mAP", "BACC", "R", "F1", "auc", "P", "hamming_loss"))}))
