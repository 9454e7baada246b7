"""Class-prototype construction + difficulty counts (kernel K2, proto.cu).

Host mirror of the inline block utils/local_training.py:973-1000 (end of stage 1) and
:1208-1249 (every stage-2 round) of the reference: label-masked feature means for the client's
annotated classes (`proto[2c]` = mean of rows with label 0, `proto[2c+1]` = label 1, rows of
other classes stay zero) and `t[c] = #{p<L or p>U}/N` for the missing classes.
One or many clients ("segments") per launch; everything stays on the device.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

from . import _cabi as cabi
from ._workspace import workspace


@dataclass
class PrototypeResult:
    proto: torch.Tensor      # [S, 2C, D] float32 (device)
    cnt: torch.Tensor        # [S, 2C] int32 (device)   num_proto of the reference
    tcnt: torch.Tensor       # [S, C] int32 (device)    confident-prediction counts
    seg_rows: list           # [S+1]

    def t(self) -> np.ndarray:
        """Per-segment t = tcnt / len(local_dataset) as float64 numpy [S, C] (synchronises)."""
        n = np.diff(np.asarray(self.seg_rows, dtype=np.int64)).astype(np.float64)
        return self.tcnt.cpu().numpy().astype(np.int64) / n[:, None]


def build_prototypes(features, labels, logits, active_classes, t_classes, L=0.3, U=0.7,
                     guard_empty=True, logits_are_probs=False, seg_rows=None) -> PrototypeResult:
    """features [N, D] fp32 CUDA, labels [N, C] fp32 (0/1), logits [N, C] fp32 or None.

    Single client: active_classes / t_classes are lists of class ids.
    Batched: seg_rows = [0, n_0, n_0+n_1, ...] and active_classes / t_classes are lists (one per
    segment) of lists.  guard_empty=True is the stage-2 behaviour (:1241-1248), False stage 1."""
    cabi.require_cuda(features, labels, logits)
    if features.dtype != torch.float32 or labels.dtype != torch.float32:
        raise TypeError("features and labels must be float32")
    features = features if features.is_contiguous() else features.contiguous()
    labels = labels.contiguous()
    logits = None if logits is None else logits.contiguous().float()
    N, D = features.shape
    C = labels.shape[1]
    if labels.shape[0] != N or (logits is not None and tuple(logits.shape) != (N, C)):
        raise ValueError("features / labels / logits row counts differ")
    if seg_rows is None:
        seg_rows = [0, N]
        active_classes, t_classes = [list(active_classes)], [list(t_classes)]
    S = len(seg_rows) - 1
    if seg_rows[-1] != N or len(active_classes) != S or len(t_classes) != S:
        raise ValueError("seg_rows / class lists inconsistent")
    dev = features.device
    proto = torch.empty(S, 2 * C, D, dtype=torch.float32, device=dev)
    cnt = torch.empty(S, 2 * C, dtype=torch.int32, device=dev)
    tcnt = torch.zeros(S, C, dtype=torch.int32, device=dev)
    lib = cabi.lib()
    with torch.cuda.device(dev):
        st = cabi.stream_ptr(dev)
        for s0 in range(0, S, cabi.MAX_SEGMENTS):
            s1 = min(S, s0 + cabi.MAX_SEGMENTS)
            r0, r1 = seg_rows[s0], seg_rows[s1]
            rows = [r - r0 for r in seg_rows[s0:s1 + 1]]
            ws_bytes = lib.fmlp_proto_ws_bytes(r1 - r0, D, C, s1 - s0)
            ws = workspace("proto", ws_bytes, dev)
            cabi.check(lib.fmlp_proto_build_f32(
                features.data_ptr() + 4 * r0 * D, D, D, labels.data_ptr() + 4 * r0 * C,
                None if logits is None else logits.data_ptr() + 4 * r0 * C, 1 if logits_are_probs else 0,
                C, s1 - s0, cabi.i64_array(rows),
                cabi.u32_array([cabi.class_mask(a) for a in active_classes[s0:s1]]),
                cabi.u32_array([cabi.class_mask(t) for t in t_classes[s0:s1]]),
                float(L), float(U), 1 if guard_empty else 0,
                proto.data_ptr() + 4 * s0 * 2 * C * D, cnt.data_ptr() + 4 * s0 * 2 * C,
                tcnt.data_ptr() + 4 * s0 * C, ws.data_ptr(), ws.numel(), st), "fmlp_proto_build_f32")
    return PrototypeResult(proto, cnt, tcnt, list(seg_rows))
