"""TEST INFRASTRUCTURE ONLY — CPU/fp32 oracle for the training-side pieces of the hot path.

  * Dice / Jaccard criterion = monai.losses.DiceLoss(include_background=True, sigmoid=True, squared_pred=True,
    jaccard=?, batch=True, reduction="mean")            src/definer.py:184-203  (MONAI 0.6.0, Appendix A: unpinned)
  * deep-supervision mean over heads                    learning/engine.py:322-330
  * Ranger2020.step (RAdam + optional GC + Lookahead)   learning/optimizer.py:136-255
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence

import torch


def dice_loss(logits: torch.Tensor, target: torch.Tensor, jaccard: bool = False, smooth_nr: float = 1e-5,
              smooth_dr: float = 1e-5) -> torch.Tensor:
    p = torch.sigmoid(logits.float())
    t = target.float()
    axes = (0, 2, 3, 4)  # batch=True: reduce over batch and space, keep channels
    inter = (t * p).sum(axes)
    ground = (t * t).sum(axes)
    pred = (p * p).sum(axes)
    denom = ground + pred
    if jaccard:
        denom = 2.0 * (denom - inter)
    f = 1.0 - (2.0 * inter + smooth_nr) / (denom + smooth_dr)
    return f.mean()


def dice_ce_loss(logits: torch.Tensor, target: torch.Tensor, lambda_dice: float = 1.0, lambda_ce: float = 1.0,
                 jaccard: bool = False) -> torch.Tensor:
    """learning/losses.py:470-595 (DiceCELoss as built at src/definer.py:204-212): the Dice term above plus
    CrossEntropyLoss(input, argmax(target, dim=1)) with mean reduction (`ce()`, losses.py:562-577)."""
    y = torch.argmax(target.float(), dim=1).long()
    ce = torch.nn.functional.cross_entropy(logits.float(), y, reduction="mean")
    return lambda_dice * dice_loss(logits, target, jaccard) + lambda_ce * ce


def deep_supervision_loss(heads: Sequence[torch.Tensor], target: torch.Tensor, jaccard: bool = False):
    """mean_k criterion(head_k, target) over main output + deep heads (engine.py:322-330)."""
    return torch.stack([dice_loss(h, target, jaccard) for h in heads]).mean()


class RangerState:
    """Per-tensor state of Ranger2020 (exp_avg, exp_avg_sq, slow_buffer, step)."""

    def __init__(self, p: torch.Tensor):
        self.step = 0
        self.exp_avg = torch.zeros_like(p, dtype=torch.float32)
        self.exp_avg_sq = torch.zeros_like(p, dtype=torch.float32)
        self.slow = p.detach().clone()


def radam_step_size(step: int, beta1: float, beta2: float, n_sma_threshold: float = 5.0):
    """(N_sma, step_size) exactly as learning/optimizer.py:205-217."""
    beta2_t = beta2 ** step
    n_sma_max = 2.0 / (1.0 - beta2) - 1.0
    n_sma = n_sma_max - 2.0 * step * beta2_t / (1.0 - beta2_t)
    if n_sma > n_sma_threshold:
        ss = math.sqrt((1 - beta2_t) * (n_sma - 4) / (n_sma_max - 4) * (n_sma - 2) / n_sma * n_sma_max /
                       (n_sma_max - 2)) / (1 - beta1 ** step)
    else:
        ss = 1.0 / (1 - beta1 ** step)
    return n_sma, ss


def ranger_step(params: List[torch.Tensor], grads: List[torch.Tensor], states: List[RangerState], lr: float,
                betas=(0.95, 0.999), eps: float = 1e-5, weight_decay: float = 0.0, alpha: float = 0.5, k: int = 6,
                n_sma_threshold: float = 5.0, use_gc: bool = False, gc_conv_only: bool = False) -> None:
    """In-place Ranger2020 update of `params` (fp32)."""
    beta1, beta2 = betas
    for p, g, st in zip(params, grads, states):
        if g is None:
            continue
        g = g.detach().float().clone()
        if use_gc and g.dim() > (3 if gc_conv_only else 1):
            g = g - g.mean(dim=tuple(range(1, g.dim())), keepdim=True)
        st.step += 1
        st.exp_avg_sq.mul_(beta2).addcmul_(g, g, value=1 - beta2)
        st.exp_avg.mul_(beta1).add_(g, alpha=1 - beta1)
        n_sma, step_size = radam_step_size(st.step, beta1, beta2, n_sma_threshold)
        if n_sma > n_sma_threshold:
            upd = st.exp_avg / (st.exp_avg_sq.sqrt() + eps)
        else:
            upd = st.exp_avg  # aliases the moving average, as `G_grad = exp_avg` does at optimizer.py:231
        if weight_decay != 0:
            if n_sma > n_sma_threshold:
                upd = upd + weight_decay * p.detach().float()
            else:
                upd.add_(p.detach().float(), alpha=weight_decay)  # in place: the decay term enters exp_avg
        p.data.add_(upd, alpha=-step_size * lr)
        if st.step % k == 0:
            st.slow.add_(p.data - st.slow, alpha=alpha)
            p.data.copy_(st.slow)
