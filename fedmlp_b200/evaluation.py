"""On-device evaluation of the global model (kernels: csrc/eval.cu).

Host mirror of utils/evaluations.py:15-73 `globaltest` and :89-140 `classtest`.  The reference
collects all probabilities on the host and runs sklearn (average_precision_score, roc_curve + auc)
and the numpy loops of utils/multilabel_metrixs.py; here the logits stay on the GPU, the per-class
sums come from `fmlp_eval_multilabel_f32` (exact pairwise counting, no sort) and only the last
arithmetic on C numbers runs on the host, written the way the reference writes it so that the
zero-division behaviour (numpy warnings, NaN / inf, skipped classes in `Precision`) is the same.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _cabi as cabi
from ._workspace import workspace


def class_statistics(scores: torch.Tensor, labels: torch.Tensor, scores_are_probs: bool = False, threshold: float = 0.5):
    """Per-class sums on the device.  scores [N, C] fp32 CUDA (logits unless scores_are_probs), labels [N, C]
    0/1.  Returns (counts int32 [C, 8] = n_pos, n_neg, n_pred, tp, tn, mismatches, 0, 0; ap_auc float64 [C, 2])."""
    cabi.require_cuda(scores, labels)
    if scores.dtype != torch.float32:
        raise TypeError("scores must be float32")
    scores = scores.contiguous()
    labels = labels.to(device=scores.device, dtype=torch.float32).contiguous()
    if scores.dim() != 2 or labels.shape != scores.shape:
        raise ValueError("scores / labels must both be [N, C]")
    N, C = scores.shape
    if N < 1:
        raise ValueError("empty test set")
    dev = scores.device
    counts = torch.zeros(C, 8, dtype=torch.int32, device=dev)
    ap_auc = torch.empty(C, 2, dtype=torch.float64, device=dev)
    lib = cabi.lib()
    ws = workspace("eval", lib.fmlp_eval_ws_bytes(N, C), dev)
    with torch.cuda.device(dev):
        cabi.check(lib.fmlp_eval_multilabel_f32(scores.data_ptr(), labels.data_ptr(), N, C, int(bool(scores_are_probs)),
                                                float(threshold), counts.data_ptr(), ap_auc.data_ptr(), ws.data_ptr(),
                                                ws.numel(), cabi.stream_ptr(dev)), "fmlp_eval_multilabel_f32")
    return counts, ap_auc


def metrics_from_class_statistics(cnt, aa, N, classid=None):
    """Last arithmetic of globaltest / classtest on the per-class sums (host, C numbers), written the way
    utils/multilabel_metrixs.py and utils/evaluations.py:41-73 write it.  cnt int [C, >=6] = n_pos, n_neg,
    n_pred, tp, tn, mismatches; aa float64 [C, 2] = AP_c, AUC_c."""
    cnt = np.asarray(cnt).astype(np.int64)
    aa = np.asarray(aa, dtype=np.float64)
    C = cnt.shape[0]
    n_pos, n_neg, n_pred, tp, tn, mism = (cnt[:, k] for k in range(6))
    with np.errstate(divide="ignore", invalid="ignore"):
        if classid is not None:                       # classtest, utils/evaluations.py:120-140
            i = int(classid)
            recall1 = tp[i] / n_pos[i]
            recall0 = tn[i] / (N - n_pos[i])
            return {"BACC": (recall0 + recall1) / 2, "R": tp[i] / n_pos[i],
                    "F1": (2 * tp[i]) / (n_pos[i] + n_pred[i]), "P": tp[i] / n_pred[i]}
        bacc = r = f1 = p = 0
        for i in range(C):                            # utils/multilabel_metrixs.py, class-wise loops
            recall1 = tp[i] / n_pos[i]
            recall0 = tn[i] / (N - n_pos[i])
            bacc += (recall0 + recall1) / 2
            r += tp[i] / n_pos[i]
            f1 += (2 * tp[i]) / (n_pos[i] + n_pred[i])
            if n_pred[i] != 0:                        # Precision skips classes without predictions, still divides by C
                p += tp[i] / n_pred[i]
        auroc = 0
        for i in range(C):                            # :60-66
            auroc += aa[i, 1]
        auroc /= C
        return {"mAP": torch.tensor([float(v) for v in aa[:, 0]]).mean(),       # torch.tensor(APs).mean(), fp32 like the reference
                "BACC": bacc / C, "R": r / C, "F1": f1 / C, "auc": auroc, "P": p / C,
                "hamming_loss": int(mism.sum()) / (N * C)}


def multilabel_metrics(scores, labels, scores_are_probs=False, classid=None):
    """The result dict of `globaltest` (classid None: mAP, BACC, R, F1, auc, P, hamming_loss) or of
    `classtest` (classid given: BACC, R, F1, P of that class) from device-resident logits."""
    counts, ap_auc = class_statistics(scores, labels, scores_are_probs)
    return metrics_from_class_statistics(counts.cpu().numpy(), ap_auc.cpu().numpy(), scores.shape[0], classid)


@torch.no_grad()
def collect_logits(net, test_dataset, args):
    """The inference loop of globaltest (:16-33) with the outputs kept on the device."""
    from torch.utils.data import DataLoader
    net.eval()
    loader = DataLoader(dataset=test_dataset, batch_size=args.batch_size * 4, shuffle=False,
                        num_workers=getattr(args, "num_workers", 4))
    outs = []
    for samples in loader:
        images = samples["image"].to(args.device)
        _, outputs = net(images)
        outs.append(outputs.detach().float())
    return torch.cat(outs)


def globaltest(net, test_dataset, args):
    """Drop-in for utils/evaluations.py:15 `globaltest(net, test_dataset, args)`."""
    logits = collect_logits(net, test_dataset, args)
    labels = torch.as_tensor(np.array(test_dataset.targets), dtype=torch.float32, device=logits.device)
    assert logits.shape[0] == len(test_dataset) and logits.shape[1] == args.n_classes
    return multilabel_metrics(logits, labels)


def classtest(net, test_dataset, args, classid):
    """Drop-in for utils/evaluations.py:89 `classtest(net, test_dataset, args, classid)`."""
    logits = collect_logits(net, test_dataset, args)
    labels = torch.as_tensor(np.array(test_dataset.targets), dtype=torch.float32, device=logits.device)
    return multilabel_metrics(logits, labels, classid=classid)
