"""Tagging fused into feature extraction (kernel: csrc/pool_tag.cu).

The reference's tagging pass (utils/local_training.py:1026-1049) runs `features, _ = net(images1)`
per batch, where the feature is the model tail flatten(adaptive_avg_pool2d(relu(fmap), 1)) of the
(patched) torchvision DenseNet / EfficientNet forward, and grows `f = torch.cat((f, features))`
until the whole dataset is in one [N, D] matrix that is then scored against the prototypes
(:1052-1058).  Here the pooling of a batch's last feature map and the scoring are ONE kernel:
`pool_tag(fmap, table, sim_out=sim[:, row0:])` returns the pooled features (for the classifier)
and has already written the batch's similarities into their columns of the [C, N] matrix.

    table = build_sim_table(Prototype, missing_classes)       # once per round
    for row0, images in batches:
        fmap = net.features(images)                            # [B, D, H, W], NCHW or channels_last
        feat = pool_tag(fmap, table, sim_out=tagger.sim, col0=row0)
    tagger.select(clean_threshold, noise_threshold)
"""
from __future__ import annotations

from dataclasses import dataclass

import torch

from . import _cabi as cabi
from .tagging import SIM_MODES

FMAP_NCHW, FMAP_NHWC = 0, 1


@dataclass
class SimTable:
    """Class vectors of one (prototype, class set, mode); valid until the prototypes change."""
    data: torch.Tensor      # device float32, fmlp_sim_table_bytes(C, D) / 4 elements
    C: int
    D: int
    classes: int            # bit mask
    mode: int


def build_sim_table(prototype: torch.Tensor, classes, mode: str = "folded", device=None) -> SimTable:
    """prototype [2C, D] (server-aggregated, reference main.py:231-234); classes: ids to score."""
    if device is None:
        device = prototype.device if prototype.is_cuda else torch.device("cuda", torch.cuda.current_device())
    prototype = prototype.to(device=device, dtype=torch.float32).contiguous()
    cabi.require_cuda(prototype)
    C, D = prototype.shape[0] // 2, prototype.shape[1]
    mask = cabi.class_mask(classes)
    lib = cabi.lib()
    data = torch.empty(lib.fmlp_sim_table_bytes(C, D) // 4, dtype=torch.float32, device=device)
    with torch.cuda.device(device):
        cabi.check(lib.fmlp_sim_table_build_f32(prototype.data_ptr(), C, D, mask, SIM_MODES[mode], data.data_ptr(),
                                                cabi.stream_ptr(device)), "fmlp_sim_table_build_f32")
    return SimTable(data, C, D, mask, SIM_MODES[mode])


def _layout_of(fmap: torch.Tensor):
    if fmap.dim() == 3:           # already [B, D, HW]
        return (fmap if fmap.is_contiguous() else fmap.contiguous()), FMAP_NCHW, fmap.shape[2]
    if fmap.dim() != 4:
        raise ValueError("feature map must be [B, D, H, W] or [B, D, HW]")
    hw = fmap.shape[2] * fmap.shape[3]
    if fmap.is_contiguous():
        return fmap, FMAP_NCHW, hw
    if fmap.is_contiguous(memory_format=torch.channels_last):
        return fmap, FMAP_NHWC, hw
    return fmap.contiguous(), FMAP_NCHW, hw


def pool_tag(fmap: torch.Tensor, table: SimTable | None = None, sim_out: torch.Tensor | None = None, col0: int = 0,
             feat_out: torch.Tensor | None = None, relu: bool = True, want_features: bool = True):
    """feat = mean_hw(relu(fmap)) [B, D]; if `table` is given, sim_out[c, col0 + b] is written for the
    table's classes in the same pass.  Inference only (the reference runs this pass under no_grad)."""
    cabi.require_cuda(fmap)
    if fmap.dtype != torch.float32:
        raise TypeError("feature map must be float32")
    if fmap.requires_grad and torch.is_grad_enabled():
        raise RuntimeError("pool_tag is forward-only; call it under torch.no_grad() like the reference's tagging pass")
    fmap, layout, HW = _layout_of(fmap.detach())
    B, D = fmap.shape[0], fmap.shape[1]
    dev = fmap.device
    feat_ptr, ld_feat = 0, 0
    if want_features or feat_out is not None:
        if feat_out is None:
            feat_out = torch.empty(B, D, dtype=torch.float32, device=dev)
        if feat_out.shape[0] < B or feat_out.shape[1] != D or feat_out.stride(1) != 1 or feat_out.dtype != torch.float32:
            raise ValueError("feat_out must be float32 [>=B, D] with unit column stride")
        feat_ptr, ld_feat = feat_out.data_ptr(), feat_out.stride(0)
    tab_ptr, sim_ptr, ld_sim, C, classes, mode = 0, 0, 0, 0, 0, cabi.SIM_FOLDED
    if table is not None and table.classes:
        if table.D != D:
            raise ValueError("table / feature-map channel mismatch")
        if sim_out is None or sim_out.dtype != torch.float32 or sim_out.dim() != 2 or sim_out.stride(1) != 1:
            raise ValueError("sim_out must be float32 [C, N]")
        if sim_out.shape[0] != table.C or col0 < 0 or col0 + B > sim_out.shape[1]:
            raise ValueError("batch does not fit in sim_out")
        tab_ptr, C, classes, mode = table.data.data_ptr(), table.C, table.classes, table.mode
        sim_ptr, ld_sim = sim_out.data_ptr() + 4 * col0, sim_out.stride(0)
    if B == 0:
        return feat_out[:0] if feat_out is not None else None
    with torch.cuda.device(dev):
        cabi.check(cabi.lib().fmlp_pool_tag_f32(fmap.data_ptr(), layout, B, D, HW, int(bool(relu)), tab_ptr, C, classes, mode,
                                                feat_ptr, ld_feat, sim_ptr, ld_sim, cabi.stream_ptr(dev)),
                   "fmlp_pool_tag_f32")
    return feat_out[:B] if feat_out is not None else None


class FusedTail(torch.nn.Module):
    """`net(x) -> (feature, logits)` (the model contract of the reference) for a backbone
    split into `features` (-> [B, D, H, W]) and `classifier` (Linear on [B, D]).

    Under torch.no_grad() — the reference's tagging and prototype passes
    (utils/local_training.py:1033, 1223) — the tail runs in the fused kernel; `tagging(x, ...)`
    additionally writes the batch's similarities.  With autograd enabled (training steps) the tail is
    the stock torch relu / adaptive_avg_pool2d chain, which is backbone work outside this library."""

    def __init__(self, features: torch.nn.Module, classifier: torch.nn.Module, relu: bool = True):
        super().__init__()
        self.features, self.classifier, self.relu = features, classifier, relu

    def forward(self, x):
        fmap = self.features(x)
        if torch.is_grad_enabled() and (fmap.requires_grad or self.training):
            out = torch.relu(fmap) if self.relu else fmap
            feat = torch.flatten(torch.nn.functional.adaptive_avg_pool2d(out, (1, 1)), 1)
        else:
            feat = pool_tag(fmap, relu=self.relu)
        return feat, self.classifier(feat)

    @torch.no_grad()
    def tagging(self, x, table: SimTable, sim_out: torch.Tensor, col0: int):
        """One batch of the tagging pass: features, logits, and sim_out[:, col0:col0+B] filled."""
        feat = pool_tag(self.features(x), table, sim_out=sim_out, col0=col0, relu=self.relu)
        return feat, self.classifier(feat)
