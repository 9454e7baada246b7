"""Pseudo-label tagging (kernels K3 / K3b / K3c: tag_sim.cu, tag_select.cu).

Host mirror of stage 2 of the reference's LocalUpdate.train_FedMLP
(utils/local_training.py:1052-1112 similarity + sign split + top-fraction selection,
:1197-1204 remaining candidates, :1456-1477 DatasetSplit_pseudo label / mask fill) for one or
many clients ("segments") resident on one GPU.

State kept on the device between rounds (the reference keeps Python lists
`self.traindata_idx` / `self.idxss`, :1025,1088-1112,1202-1204):
    tag[c, n] = 0  row n is still a candidate for missing class c          (in `idxss`)
              = 1  confidently negative ("clean" list,  traindata_idx[2i])
              = 2  confidently positive ("noise" list,  traindata_idx[2i+1])
The Python lists are materialised lazily from the per-round selections when asked for.
"""
from __future__ import annotations

import math

import torch

from . import _cabi as cabi
from ._workspace import workspace

SIM_MODES = {"pair": cabi.SIM_PAIR, "folded": cabi.SIM_FOLDED}


def tag_similarity(features, prototype, missing_classes, seg_rows=None, out=None, mode="pair"):
    """sim[c, n] = cos(f_n, P[2c]) - cos(f_n, P[2c+1]) for the missing classes, one pass over
    `features` (reference :1052-1058 + CosineSimilarityFast :1417-1435).

    features [N, D] fp32 CUDA; prototype [2C, D] fp32; missing_classes: list of class ids
    (single client) or list of lists with seg_rows.  Returns sim [C, N] (rows of classes that are
    not missing are left as they were in `out`, NaN in a fresh buffer)."""
    cabi.require_cuda(features, prototype)
    if features.dtype != torch.float32:
        raise TypeError("features must be float32")
    features = features if features.is_contiguous() else features.contiguous()
    prototype = prototype.to(device=features.device, dtype=torch.float32).contiguous()
    N, D = features.shape
    C = prototype.shape[0] // 2
    if prototype.shape[1] != D:
        raise ValueError("prototype / feature dimension mismatch")
    if seg_rows is None:
        seg_rows, missing_classes = [0, N], [list(missing_classes)]
    S = len(seg_rows) - 1
    dev = features.device
    if out is None:
        out = torch.full((C, N), float("nan"), dtype=torch.float32, device=dev)
    lib = cabi.lib()
    with torch.cuda.device(dev):
        st = cabi.stream_ptr(dev)
        ws = workspace("sim", lib.fmlp_tag_sim_ws_bytes(C, D), dev)      # class-vector table of the pre-kernel
        for s0 in range(0, S, cabi.MAX_SEGMENTS):
            s1 = min(S, s0 + cabi.MAX_SEGMENTS)
            r0 = seg_rows[s0]
            rows = [r - r0 for r in seg_rows[s0:s1 + 1]]
            cabi.check(lib.fmlp_tag_sim_f32(
                features.data_ptr() + 4 * r0 * D, D, D, prototype.data_ptr(), C, s1 - s0,
                cabi.i64_array(rows), cabi.u32_array([cabi.class_mask(m) for m in missing_classes[s0:s1]]),
                out.data_ptr() + 4 * r0, N, SIM_MODES[mode], ws.data_ptr(), ws.numel(), st), "fmlp_tag_sim_f32")
    return out


class TagBatch:
    """Tagging state of S clients whose rows are stored back to back in one [N_total, D] matrix."""

    def __init__(self, seg_rows, n_classes, active_classes, missing_classes, dataset_idx=None, device=None):
        self.seg_rows = [int(r) for r in seg_rows]
        self.S = len(self.seg_rows) - 1
        self.C = int(n_classes)
        self.N = self.seg_rows[-1]
        self.active = [list(map(int, a)) for a in active_classes]
        self.missing = [list(map(int, m)) for m in missing_classes]
        if len(self.active) != self.S or len(self.missing) != self.S:
            raise ValueError("one active / missing class list per segment")
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device())
        self.device = torch.device(device)
        self.tag = torch.zeros(self.C, max(self.N, 1), dtype=torch.uint8, device=self.device)
        self.sim = torch.full((self.C, max(self.N, 1)), float("nan"), dtype=torch.float32, device=self.device)
        if dataset_idx is None:
            # local row number within the segment
            dataset_idx = torch.cat([torch.arange(self.seg_rows[s + 1] - self.seg_rows[s]) for s in range(self.S)]) \
                if self.S else torch.zeros(0, dtype=torch.int64)
        self.dataset_idx = torch.as_tensor(dataset_idx).to(self.device, dtype=torch.int64)
        self._history = []   # per step: (counts [S,C,4] dev, sel [S,C,2,cap] dev, cap)
        self._lists = None

    # ---------------------------------------------------------------------------------------
    def step(self, features, prototype, clean_frac=0.005, noise_frac=0.01, mode="pair"):
        """One tagging pass (reference :1052-1112): similarity over all rows, selection among
        the rows that are still candidates, tag-state update.  Returns (counts, sel, cap) device
        tensors: counts[s, c] = (n_clean, n_noise, m, k); sel[s, c, side, :m|k] = global rows in
        rank order."""
        self.similarity(features, prototype, mode)
        return self.select(clean_frac, noise_frac)

    def similarity(self, features, prototype, mode="pair"):
        """Fill self.sim [C, N] for every segment's missing classes (reference :1052-1058)."""
        if features.shape[0] != self.N:
            raise ValueError("feature rows != rows of the tag batch")
        tag_similarity(features, prototype, self.missing, self.seg_rows, out=self.sim, mode=mode)
        return self.sim

    def select(self, clean_frac=0.005, noise_frac=0.01):
        """Sign split + top-fraction selection on self.sim among the candidate rows, and tag-state
        update (reference :1061-1112, utils/utils.py:24-35)."""
        max_rows = max((self.seg_rows[s + 1] - self.seg_rows[s] for s in range(self.S)), default=0)
        cap = max(1, int(math.floor(max(clean_frac, noise_frac, 0.0) * max_rows)) + 1)
        cap = min(cap, max(max_rows, 1))
        counts = torch.zeros(self.S, self.C, 4, dtype=torch.int32, device=self.device)
        self.remaining_count = torch.zeros(self.S, self.C, dtype=torch.int32, device=self.device)
        sel = torch.full((self.S, self.C, 2, cap), -1, dtype=torch.int32, device=self.device)
        lib = cabi.lib()
        with torch.cuda.device(self.device):
            st = cabi.stream_ptr(self.device)
            for s0 in range(0, self.S, cabi.MAX_SEGMENTS):
                s1 = min(self.S, s0 + cabi.MAX_SEGMENTS)
                r0 = self.seg_rows[s0]
                rows = [r - r0 for r in self.seg_rows[s0:s1 + 1]]
                ws_bytes = lib.fmlp_tag_select_ws_bytes(s1 - s0, self.C, cap)
                ws = workspace("select", ws_bytes, self.device)
                cabi.check(lib.fmlp_tag_select(
                    self.sim.data_ptr() + 4 * r0, self.sim.shape[1], self.tag.data_ptr() + r0, self.tag.shape[1],
                    self.C, s1 - s0, cabi.i64_array(rows),
                    cabi.u32_array([cabi.class_mask(m) for m in self.missing[s0:s1]]),
                    float(clean_frac), float(noise_frac), counts.data_ptr() + 4 * s0 * self.C * 4,
                    self.remaining_count.data_ptr() + 4 * s0 * self.C,
                    sel.data_ptr() + 4 * s0 * self.C * 2 * cap, cap, ws.data_ptr(), ws.numel(), st),
                    "fmlp_tag_select")
                if r0:  # rows reported by a later group are relative to that group's first row
                    blk = sel[s0:s1]
                    blk[blk >= 0] += r0
        self._history.append((counts, sel, cap))
        self._lists = None
        return counts, sel, cap

    # ---------------------------------------------------------------------------------------
    def fill(self, labels):
        """Label / mask fill for all rows (reference DatasetSplit_pseudo.__getitem__ :1456-1477
        and sup_cls :1173).  labels [N, C] fp32 original targets.  Returns (y, distill, sup)."""
        cabi.require_cuda(labels)
        labels = labels.to(torch.float32).contiguous()
        if tuple(labels.shape) != (self.N, self.C):
            raise ValueError("labels must be [N_total, C]")
        y = torch.empty_like(labels)
        distill = torch.empty_like(labels)
        sup = torch.empty_like(labels)
        lib = cabi.lib()
        with torch.cuda.device(self.device):
            st = cabi.stream_ptr(self.device)
            for s0 in range(0, self.S, cabi.MAX_SEGMENTS):
                s1 = min(self.S, s0 + cabi.MAX_SEGMENTS)
                r0 = self.seg_rows[s0]
                rows = [r - r0 for r in self.seg_rows[s0:s1 + 1]]
                o = 4 * r0 * self.C
                cabi.check(lib.fmlp_mask_fill(
                    labels.data_ptr() + o, self.tag.data_ptr() + r0, self.tag.shape[1], self.C, s1 - s0,
                    cabi.i64_array(rows), cabi.u32_array([cabi.class_mask(a) for a in self.active[s0:s1]]),
                    cabi.u32_array([cabi.class_mask(m) for m in self.missing[s0:s1]]),
                    y.data_ptr() + o, distill.data_ptr() + o, sup.data_ptr() + o, st), "fmlp_mask_fill")
        return y, distill, sup

    # ---------------------------------------------------------------------------------------
    def traindata_idx(self, s=0):
        """The reference's `self.traindata_idx` of client s: [clean_c0, noise_c0, clean_c1, ...]
        in the client's missing-class order, dataset indices as Python floats in pick order
        (first stage-2 round appends, later rounds extend; :1088-1089,1111-1112).  Synchronises."""
        if self._lists is None:
            lists = [[[] for _ in range(2 * len(self.missing[q]))] for q in range(self.S)]
            ids = self.dataset_idx.cpu().numpy()
            for counts, sel, cap in self._history:
                cnt = counts.cpu().numpy()
                rows = sel.cpu().numpy()
                for q in range(self.S):
                    for i, c in enumerate(self.missing[q]):
                        m, k = int(cnt[q, c, 2]), int(cnt[q, c, 3])
                        lists[q][2 * i].extend(float(v) for v in ids[rows[q, c, 0, :m]])
                        lists[q][2 * i + 1].extend(float(v) for v in ids[rows[q, c, 1, :k]])
            self._lists = lists
        return self._lists[s]

    def remaining(self, s=0):
        """The reference's `self.idxss` of client s (:1197-1204): per missing class the dataset
        indices in neither list, as a sorted list (the reference's order is set-iteration order)."""
        r0, r1 = self.seg_rows[s], self.seg_rows[s + 1]
        tag = self.tag[:, r0:r1].cpu().numpy()
        ids = self.dataset_idx[r0:r1].cpu().numpy()
        return [sorted(int(v) for v in ids[tag[c] == 0]) for c in self.missing[s]]

    def class_num_noise(self, s=0):
        """len(traindata_idx[2i+1]) per missing class (reference :1117-1120 class_num_list update)."""
        r0, r1 = self.seg_rows[s], self.seg_rows[s + 1]
        t = self.tag[:, r0:r1]
        return [int((t[c] == 2).sum()) for c in self.missing[s]]
