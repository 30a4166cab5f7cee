"""Dice / Jaccard criterion of the reference training step on the fused CUDA kernels.

Drop-in for ``monai.losses.DiceLoss(include_background=True, sigmoid=True, squared_pred=True, jaccard=?, batch=True,
reduction="mean")`` exactly as ``src/definer.py:184-203`` builds it (``--criterion dice`` / ``jaccard``):
``forward(input[N, K, D, H, W] logits, target[N, K, D, H, W]) -> scalar`` with autograd support.  One reduction
kernel (sigmoid + the three per-channel sums) and one backward kernel, instead of ~10 elementwise/reduction passes.
"""
from __future__ import annotations

import torch
from torch import nn

from . import ops


class _DiceFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target, jaccard, smooth_nr, smooth_dr, lambda_dice=1.0, lambda_ce=0.0):
        if not logits.is_cuda:
            raise RuntimeError("brats21_b200.DiceLoss runs on CUDA only (no CPU fallback)")
        x = logits.detach().to(torch.float32).contiguous()
        t = target.detach().to(torch.float32).contiguous()
        if x.shape != t.shape:
            raise AssertionError(f"ground truth has differing shape ({tuple(t.shape)}) from input ({tuple(x.shape)})")
        loss = torch.zeros((1,), dtype=torch.float32, device=x.device)
        coef = ops.dice_fwd(x, t, loss, jaccard=jaccard, weight=lambda_dice, smooth_nr=smooth_nr, smooth_dr=smooth_dr)
        if lambda_dice != 1.0:
            coef = coef * lambda_dice
        if lambda_ce != 0.0:
            ops.ce_fwd(x, t, loss, weight=lambda_ce)
        ctx.save_for_backward(x, t, coef)
        ctx.lambda_ce = float(lambda_ce)
        return loss[0]

    @staticmethod
    def backward(ctx, gout):
        x, t, coef = ctx.saved_tensors
        g = gout.detach().to(torch.float32).reshape(1).contiguous()
        return ops.dice_bwd(x, t, coef, gout=g, ce_weight=ctx.lambda_ce), None, None, None, None, None, None


class DiceLoss(nn.Module):
    def __init__(self, include_background: bool = True, sigmoid: bool = True, squared_pred: bool = True,
                 jaccard: bool = False, batch: bool = True, reduction: str = "mean", smooth_nr: float = 1e-5,
                 smooth_dr: float = 1e-5):
        super().__init__()
        if not (include_background and sigmoid and squared_pred and batch and reduction == "mean"):
            raise NotImplementedError("only the configuration the reference trains with is on the accelerated path: "
                                      "include_background, sigmoid, squared_pred, batch, reduction='mean'")
        self.jaccard, self.smooth_nr, self.smooth_dr = bool(jaccard), float(smooth_nr), float(smooth_dr)

    def forward(self, input: torch.Tensor, target: torch.Tensor) -> torch.Tensor:  # noqa: A002 (reference name)
        return _DiceFn.apply(input, target, self.jaccard, self.smooth_nr, self.smooth_dr)



class DiceCELoss(nn.Module):
    """Drop-in for ``learning.losses.DiceCELoss`` (learning/losses.py:470-595) as ``src/definer.py:204-212`` builds it
    for ``--criterion dice_ce``: ``lambda_dice * DiceLoss(sigmoid, squared_pred, batch) + lambda_ce *
    CrossEntropyLoss(input, argmax(target, dim=1))``, fused into the Dice kernels (one extra reduction pass forward,
    none backward)."""

    def __init__(self, include_background: bool = True, to_onehot_y: bool = False, sigmoid: bool = False,
                 softmax: bool = False, other_act=None, squared_pred: bool = False, jaccard: bool = False,
                 reduction: str = "mean", smooth_nr: float = 1e-5, smooth_dr: float = 1e-5, batch: bool = False,
                 ce_weight=None, lambda_dice: float = 1.0, lambda_ce: float = 1.0):
        super().__init__()
        if not (include_background and sigmoid and squared_pred and batch and reduction == "mean") or to_onehot_y \
                or softmax or other_act is not None or ce_weight is not None:
            raise NotImplementedError("only the configuration the reference trains with is on the accelerated path: "
                                      "include_background, sigmoid, squared_pred, batch, reduction='mean'")
        if lambda_dice < 0.0:
            raise ValueError("lambda_dice should be no less than 0.0.")
        if lambda_ce < 0.0:
            raise ValueError("lambda_ce should be no less than 0.0.")
        self.jaccard, self.smooth_nr, self.smooth_dr = bool(jaccard), float(smooth_nr), float(smooth_dr)
        self.lambda_dice, self.lambda_ce = float(lambda_dice), float(lambda_ce)

    def forward(self, input: torch.Tensor, target: torch.Tensor) -> torch.Tensor:  # noqa: A002 (reference name)
        if len(input.shape) != len(target.shape):
            raise ValueError("the number of dimensions for input and target should be the same.")
        if target.shape[1] != input.shape[1]:
            raise NotImplementedError("DiceCELoss: label-map targets (B1HWD) are not on the accelerated path; pass the "
                                      "multi-channel target the BraTS loaders produce")
        return _DiceFn.apply(input, target, self.jaccard, self.smooth_nr, self.smooth_dr, self.lambda_dice,
                             self.lambda_ce)
