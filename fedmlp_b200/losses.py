"""FedMLP training losses (kernel K4, loss.cu) behind torch.autograd.Function.

Loss entry points of the reference's LocalUpdate.train_FedMLP as callables on [B, C] logits that
return a scalar with autograd:
    fedmlp_stage1_loss  — utils/local_training.py:933-963  (masked BCE on the annotated classes +
                          MSE-on-sigmoid consistency with the frozen global model on the missing ones)
    fedmlp_stage2_loss  — utils/local_training.py:1171-1188 (BCE on supervised entries / sum(sup);
                          variant 'sup_dis' is the commented alternative at :1187)
One fused launch computes the loss AND d loss / d logits; backward() only multiplies by the
incoming gradient (read on the device, no host sync).
"""
from __future__ import annotations

import torch

from . import _cabi as cabi
from ._workspace import workspace

LOSS2_VARIANTS = {"sup": cabi.LOSS2_SUP, "sup_dis": cabi.LOSS2_SUP_DIS}


def _prep(t, ref=None):
    cabi.require_cuda(t)
    if t.dtype != torch.float32:
        raise TypeError("loss inputs must be float32")
    t = t.contiguous()
    if ref is not None and t.shape != ref.shape:
        raise ValueError(f"shape mismatch {tuple(t.shape)} vs {tuple(ref.shape)}")
    return t


def _scale_by(dz, grad_out):
    """dz *= grad_out (a 0-dim device tensor), in place, without a host sync."""
    g = grad_out.to(device=dz.device, dtype=torch.float32).reshape(1).contiguous()
    with torch.cuda.device(dz.device):
        cabi.check(cabi.lib().fmlp_scale_f32(dz.data_ptr(), dz.numel(), g.data_ptr(), cabi.stream_ptr(dz.device)),
                   "fmlp_scale_f32")
    return dz


def launch_stage1(z1, z2, z3, z4, y, active_mask, missing_mask, batch_size, loss_out, dz1_out, dz2_out):
    """Raw launch of fmlp_loss_stage1_f32 into caller-provided outputs (all contiguous fp32 CUDA)."""
    B, C = z1.shape
    dev = z1.device
    lib = cabi.lib()
    with torch.cuda.device(dev):
        ws = workspace("loss", lib.fmlp_loss_ws_bytes(B, C), dev)
        cabi.check(lib.fmlp_loss_stage1_f32(z1.data_ptr(), z2.data_ptr(), z3.data_ptr(), z4.data_ptr(),
                                            y.data_ptr(), B, C, active_mask, missing_mask, int(batch_size),
                                            loss_out.data_ptr(), dz1_out.data_ptr(), dz2_out.data_ptr(), ws.data_ptr(),
                                            ws.numel(), cabi.stream_ptr(dev)), "fmlp_loss_stage1_f32")


def launch_stage2(z, zg, y, distill, variant, loss_out, dz_out):
    """Raw launch of fmlp_loss_stage2_f32 into caller-provided outputs (all contiguous fp32 CUDA)."""
    B, C = z.shape
    dev = z.device
    lib = cabi.lib()
    with torch.cuda.device(dev):
        ws = workspace("loss", lib.fmlp_loss_ws_bytes(B, C), dev)
        cabi.check(lib.fmlp_loss_stage2_f32(z.data_ptr(), None if zg is None else zg.data_ptr(), y.data_ptr(),
                                            distill.data_ptr(), B, C, variant, loss_out.data_ptr(), dz_out.data_ptr(),
                                            ws.data_ptr(), ws.numel(), cabi.stream_ptr(dev)), "fmlp_loss_stage2_f32")


def launch_stage2_seg(z, zg, y, distill, seg_rows, variant, loss_out, dz_out, seg_class_distill=None):
    """Raw launch of fmlp_loss_stage2_seg_f32: S clients stored back to back, one loss per client."""
    N, C = z.shape
    S = len(seg_rows) - 1
    if S > cabi.MAX_SEGMENTS:
        raise ValueError(f"at most {cabi.MAX_SEGMENTS} segments per launch")
    dev = z.device
    lib = cabi.lib()
    with torch.cuda.device(dev):
        ws = workspace("loss", lib.fmlp_loss_ws_bytes(N, C), dev)
        cabi.check(lib.fmlp_loss_stage2_seg_f32(z.data_ptr(), None if zg is None else zg.data_ptr(), y.data_ptr(),
                                                distill.data_ptr(), C, S, cabi.i64_array(seg_rows), variant,
                                                None if seg_class_distill is None else seg_class_distill.data_ptr(),
                                                loss_out.data_ptr(), dz_out.data_ptr(), ws.data_ptr(), ws.numel(),
                                                cabi.stream_ptr(dev)), "fmlp_loss_stage2_seg_f32")


class _Stage1(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z1, z2, z3, z4, y, active_mask, missing_mask, batch_size):
        z1 = _prep(z1); z2 = _prep(z2, z1); z3 = _prep(z3, z1); z4 = _prep(z4, z1); y = _prep(y, z1)
        B, C = z1.shape
        dev = z1.device
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        dz1 = torch.empty_like(z1)
        dz2 = torch.empty_like(z2)
        launch_stage1(z1, z2, z3, z4, y, active_mask, missing_mask, batch_size, loss, dz1, dz2)
        ctx.save_for_backward(dz1, dz2)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, grad_out):
        dz1, dz2 = ctx.saved_tensors
        return _scale_by(dz1.clone(), grad_out), _scale_by(dz2.clone(), grad_out), None, None, None, None, None, None


class _Stage2(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, zg, y, distill, variant):
        z = _prep(z); y = _prep(y, z); distill = _prep(distill, z)
        zg = None if zg is None else _prep(zg, z)
        B, C = z.shape
        dev = z.device
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        dz = torch.empty_like(z)
        launch_stage2(z, zg, y, distill, variant, loss, dz)
        ctx.save_for_backward(dz)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, grad_out):
        (dz,) = ctx.saved_tensors
        return _scale_by(dz.clone(), grad_out), None, None, None, None


def fedmlp_stage1_loss(logits1, logits2, logits_glob1, logits_glob2, labels, active_classes, missing_classes,
                       batch_size):
    """Stage-1 loss of the reference (:933-963).  logits1/2: student logits of the two augmented
    views (grad flows), logits_glob1/2: frozen global model, labels [B, C] 0/1 floats.
    batch_size is args.batch_size — the reference divides by it, not by the actual B (:956-959)."""
    return _Stage1.apply(logits1, logits2, logits_glob1.detach(), logits_glob2.detach(), labels,
                         cabi.class_mask(active_classes), cabi.class_mask(missing_classes), int(batch_size))


def fedmlp_stage2_loss(logits, logits_glob, labels, distill_cls, variant="sup"):
    """Stage-2 loss of the reference (:1171-1188).  distill_cls [B, C] is DatasetSplit_pseudo's
    mask (1 = no pseudo label, not supervised); sup_cls = ~distill_cls."""
    zg = None if logits_glob is None else logits_glob.detach()
    return _Stage2.apply(logits, zg, labels, distill_cls, LOSS2_VARIANTS[variant])


def fused_loss_and_grad_stage1(z1, z2, z3, z4, y, active_classes, missing_classes, batch_size):
    """Direct (non-autograd) access to the fused kernel: returns (loss[1], dz1, dz2)."""
    with torch.no_grad():
        z1r = z1.detach().requires_grad_(True)
        z2r = z2.detach().requires_grad_(True)
    with torch.enable_grad():
        loss = fedmlp_stage1_loss(z1r, z2r, z3, z4, y, active_classes, missing_classes, batch_size)
    dz1, dz2 = loss.grad_fn.saved_tensors
    return loss.detach(), dz1, dz2


def fused_loss_and_grad_stage2(z, zg, y, distill, variant="sup"):
    with torch.enable_grad():
        zr = z.detach().requires_grad_(True)
        loss = fedmlp_stage2_loss(zr, zg, y, distill, variant)
    (dz,) = loss.grad_fn.saved_tensors
    return loss.detach(), dz
