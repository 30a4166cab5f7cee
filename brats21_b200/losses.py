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
    def forward(ctx, logits, target, jaccard, smooth_nr, smooth_dr):
        if not logits.is_cuda:
            raise RuntimeError("brats21_b200.DiceLoss runs on CUDA only (no CPU fallback)")
        x = logits.detach().to(torch.float32).contiguous()
        t = target.detach().to(torch.float32).contiguous()
        if x.shape != t.shape:
            raise AssertionError(f"ground truth has differing shape ({tuple(t.shape)}) from input ({tuple(x.shape)})")
        loss = torch.zeros((1,), dtype=torch.float32, device=x.device)
        coef = ops.dice_fwd(x, t, loss, jaccard=jaccard, smooth_nr=smooth_nr, smooth_dr=smooth_dr)
        ctx.save_for_backward(x, t, coef)
        return loss[0]

    @staticmethod
    def backward(ctx, gout):
        x, t, coef = ctx.saved_tensors
        g = gout.detach().to(torch.float32).reshape(1).contiguous()
        return ops.dice_bwd(x, t, coef, gout=g), None, None, None, None


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
