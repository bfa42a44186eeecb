"""Loss / decode / metrics with the reference's call surface (src/utils/loss.py), CUDA-backed.

Inputs may be CUDA tensors (nothing leaves the device), or -- exactly what the reference accepts
-- CPU tensors / numpy arrays, which are shipped to the current CUDA device, computed there and
shipped back.  There is no CPU implementation behind any of these.
"""
from __future__ import annotations

import sys

import numpy as np
import torch
import torch.nn as nn

from .. import ops
from ..ops import LighthandError


def _to_cuda(x, dtype=torch.float32):
    ops.require_cuda()
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(np.ascontiguousarray(x))
    if not isinstance(x, torch.Tensor):
        x = torch.as_tensor(x)
    if x.dtype != dtype:
        x = x.to(dtype)
    return x if x.is_cuda else x.cuda()


class _JointsMSE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, output, target, weight):
        need_grad = output.requires_grad
        loss, grad = ops.jmse_loss(output.detach(), target.detach(), weight, want_grad=need_grad)
        ctx.grad = grad
        return loss

    @staticmethod
    def backward(ctx, grad_output):
        g = ctx.grad
        ctx.grad = None
        if g is None:
            return None, None, None
        return g.mul_(grad_output), None, None


class JointsMSELoss(nn.Module):
    """Reference: src/utils/loss.py:306-325.  ``loss = sum_j 0.5*MSE(pred_j, gt_j) / J``; forward and
    backward are ONE fused kernel (the gradient is emitted while the loss is reduced).

    Like the reference, ``forward`` ignores ``target_weight`` and ``use_target_weight`` (Appendix B.1 of
    SURVEY.md).  ``apply_target_weight=True`` (not in the reference) enables the upstream-HRNet
    masking ``pred*w, gt*w``."""

    def __init__(self, use_target_weight, apply_target_weight=False):
        super().__init__()
        self.use_target_weight = use_target_weight
        self.apply_target_weight = apply_target_weight

    def forward(self, output, target, target_weight=None):
        if not (isinstance(output, torch.Tensor) and output.is_cuda):
            raise LighthandError("JointsMSELoss: output must be a CUDA tensor (no CPU fallback)")
        if output.dim() == 3:
            raise LighthandError("JointsMSELoss: expected [N,J,H,W]")
        target = _to_cuda(target).to(output.device)
        w = None
        if self.apply_target_weight and target_weight is not None:
            w = _to_cuda(target_weight).to(output.device)
        if output.dtype != torch.float32:
            output = output.float()
        return _JointsMSE.apply(output.contiguous(), target.contiguous(), w)


def get_max_preds(batch_heatmaps):
    """Reference: src/utils/loss.py:327-355.  numpy in -> numpy out (as the reference asserts);
    CUDA tensor in -> CUDA tensors out (no host round trip)."""
    if isinstance(batch_heatmaps, torch.Tensor) and batch_heatmaps.is_cuda:
        assert batch_heatmaps.dim() == 4, 'batch_images should be 4-ndim'
        coords, maxvals, _ = ops.decode(batch_heatmaps.float(), 1.0)
        return coords, maxvals
    assert isinstance(batch_heatmaps, np.ndarray), 'batch_heatmaps should be numpy.ndarray'
    assert batch_heatmaps.ndim == 4, 'batch_images should be 4-ndim'
    coords, maxvals, _ = ops.decode(_to_cuda(batch_heatmaps), 1.0)
    return coords.cpu().numpy(), maxvals.cpu().numpy().astype(batch_heatmaps.dtype, copy=False)


def get_final_preds(batch_heatmaps, out_scale=4.0, subpixel=True, inv_affine=None):
    """Decode + the two sub-steps the reference's configs name but never run (POST_PROCESS quarter-pixel
    shift, inverse crop affine).  CUDA tensors in/out."""
    hm = _to_cuda(batch_heatmaps)
    if inv_affine is not None:
        inv_affine = _to_cuda(inv_affine, torch.float64).to(hm.device)
    return ops.decode(hm, out_scale, subpixel, inv_affine)


def _pck_counts(pred_2d, gt_2d, T, threshold, visible):
    pred = _to_cuda(pred_2d.detach() if isinstance(pred_2d, torch.Tensor) else pred_2d)
    gt = _to_cuda(gt_2d.detach() if isinstance(gt_2d, torch.Tensor) else gt_2d).to(pred.device)
    counts, sums = ops.new_metric_state(pred.device)
    ops.pck_epe_batch(pred, gt, T, threshold, visible, counts, sums)
    c = counts.cpu().tolist()
    return c, sums.cpu().tolist()


def PCK_2d_loss(pred_2d, gt_2d, T=0.1, threshold='proportion'):
    """Reference: src/utils/loss.py:116-148."""
    c, _ = _pck_counts(pred_2d, gt_2d, T, threshold, False)
    return float((c[1] - c[0]) / c[1])


def PCK_2d_loss_visible(pred_2d, gt_2d, T=0.1, threshold='proportion'):
    """Reference: src/utils/loss.py:83-114."""
    c, _ = _pck_counts(pred_2d, gt_2d, T, threshold, True)
    return (c[1] - c[0]) / (c[1] + sys.float_info.epsilon)


def PCK_2d_loss_list(pred_2d, gt_2d, T_list=None, threshold='proportion', auc_t=None):
    """Reference: src/utils/loss.py:150-202."""
    pck_list = [[T, PCK_2d_loss(pred_2d, gt_2d, T, threshold) * 100] for T in T_list]
    auc_list = [PCK_2d_loss(pred_2d, gt_2d, T, threshold) * 100 for T in (auc_t or [])]
    return pck_list, auc_list


def EPE_train(pred_2d_joints, gt_2d_joint):
    """Reference: src/utils/loss.py:50-67 -> ((sum, count), {}).  Joints 1..J-2 (the reference's
    ``range(1, len(distance))`` drops the last joint).  The per-joint dict the reference also returns
    is unused by every caller and is returned empty."""
    c, s = _pck_counts(pred_2d_joints, gt_2d_joint, 0.0, 'mm', False)
    return (s[0], float(c[2])), {}


def EPE(pred_2d_joints, gt_2d_joint):
    """Reference: src/utils/loss.py:28-47 (skips joints whose visibility flag is 0)."""
    c, s = _pck_counts(pred_2d_joints, gt_2d_joint, 0.0, 'mm', True)
    return (s[0], float(c[2])), {}
