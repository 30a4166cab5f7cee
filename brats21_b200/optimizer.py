"""Ranger2020 (RAdam + Lookahead) with the reference's constructor and state layout, stepping ALL parameter tensors in
one fused CUDA launch (learning/optimizer.py:62-255 runs a Python loop with ~20 small kernels per tensor).

State per parameter, as in the reference: ``step``, ``exp_avg``, ``exp_avg_sq``, ``slow_buffer``.  The scalars that
depend only on the step count (N_sma, step size, look-ahead phase) are computed on the host exactly as
``optimizer.py:205-217`` does, including its 10-slot buffer semantics (all tensors share the step count).
``grad_scale`` multiplies every gradient inside the kernel (1/world_size for the data-parallel mean, or a loss
scale), so no separate un-scaling pass is needed.
"""
from __future__ import annotations

import math

import torch
from torch.optim.optimizer import Optimizer

from . import _lib
from ._lib import call, ptr, stream_ptr


class Ranger2020(Optimizer):
    def __init__(self, params, lr=1e-3, alpha=0.5, k=6, N_sma_threshhold=5, betas=(.95, 0.999), eps=1e-5,
                 weight_decay=0, use_gc=True, use_gcnorm=False, normloss=False, normloss_factor=1e-4,
                 gc_conv_only=False, gc_loc=True, normloss_active=None):
        """Same positional order, keyword names and defaults as learning/optimizer.py:62-78 (``use_gc=True``,
        ``normloss=``); ``normloss_active`` is kept as an alias of ``normloss``."""
        if normloss_active is not None:
            normloss = normloss_active
        if not 0.0 <= alpha <= 1.0:
            raise ValueError(f'Invalid slow update rate: {alpha}')
        if not 1 <= k:
            raise ValueError(f'Invalid lookahead steps: {k}')
        if not lr > 0:
            raise ValueError(f'Invalid Learning Rate: {lr}')
        if not eps > 0:
            raise ValueError(f'Invalid eps: {eps}')
        if use_gcnorm or normloss or (use_gc and not gc_loc):
            raise NotImplementedError("gradient normalisation / norm loss / post-moment centralisation are off in "
                                      "the reference recipe (README.md:103-121) and not on the fused path")
        defaults = dict(lr=lr, alpha=alpha, k=k, betas=betas, N_sma_threshhold=N_sma_threshhold, eps=eps,
                        weight_decay=weight_decay)
        super().__init__(params, defaults)
        self.N_sma_threshhold, self.alpha, self.k = N_sma_threshhold, alpha, k
        self.use_gc, self.gc_conv_only, self.gc_loc, self.use_gcnorm = use_gc, gc_conv_only, gc_loc, use_gcnorm
        self.normloss_active, self.normloss_factor, self.eps = normloss, normloss_factor, eps
        self.grad_scale = 1.0
        self._tables = {}

    def _table(self, gi, plist):
        # the device table stores raw pointers to the parameter, its gradient AND its three state tensors: all of
        # them are part of the key (Optimizer.load_state_dict replaces the state tensors without touching p / p.grad)
        key = (gi, self.use_gc, self.gc_conv_only,
               tuple((p.data_ptr(), p.grad.data_ptr(), self.state[p]["exp_avg"].data_ptr(),
                      self.state[p]["exp_avg_sq"].data_ptr(), self.state[p]["slow_buffer"].data_ptr()) for p in plist))
        hit = self._tables.get(gi)
        if hit is not None and hit[0] == key:
            return hit[1:]
        chunk = _lib.load().b21_ranger_chunk()
        rows, chunks = [], []
        for ti, p in enumerate(plist):
            st = self.state[p]
            rows.append([p.data_ptr(), p.grad.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(),
                         st["slow_buffer"].data_ptr(), p.numel()])
            chunks += [[ti, off] for off in range(0, p.numel(), chunk)]
        dev = plist[0].device
        table = torch.tensor(rows, dtype=torch.int64, device=dev)
        ctab = torch.tensor(chunks, dtype=torch.int32, device=dev)
        gc_rows = None
        if self.use_gc:  # centralized_gradient (optimizer.py:11-20): one row per dim-0 slice of qualifying tensors
            min_dim = 3 if self.gc_conv_only else 1
            rr = []
            for p in plist:
                if p.dim() > min_dim:
                    rl = p.numel() // p.shape[0]
                    rr += [[p.grad.data_ptr() + 4 * rl * i, rl] for i in range(p.shape[0])]
            gc_rows = torch.tensor(rr, dtype=torch.int64, device=dev) if rr else None
        self._tables[gi] = (key, table, ctab, gc_rows)
        return table, ctab, gc_rows

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        for gi, group in enumerate(self.param_groups):
            plist = [p for p in group["params"] if p.grad is not None]
            if not plist:
                continue
            for p in plist:
                if not p.is_cuda:
                    raise RuntimeError("brats21_b200.Ranger2020 runs on CUDA only (no CPU fallback)")
                if p.dtype != torch.float32 or p.grad.dtype != torch.float32 or not p.is_contiguous() \
                        or not p.grad.is_contiguous():
                    raise RuntimeError("Ranger2020 fused step needs contiguous fp32 parameters and gradients")
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p)
                    st["exp_avg_sq"] = torch.zeros_like(p)
                    st["slow_buffer"] = p.detach().clone()
                st["step"] += 1
            step = self.state[plist[0]]["step"]
            if any(self.state[p]["step"] != step for p in plist):
                raise RuntimeError("fused Ranger2020 step expects all parameters of a group to share the step count")
            beta1, beta2 = group["betas"]
            beta2_t = beta2 ** step
            n_sma_max = 2 / (1 - beta2) - 1
            n_sma = n_sma_max - 2 * step * beta2_t / (1 - beta2_t)
            rect = n_sma > self.N_sma_threshhold
            if rect:
                step_size = math.sqrt((1 - beta2_t) * (n_sma - 4) / (n_sma_max - 4) * (n_sma - 2) / n_sma * n_sma_max /
                                      (n_sma_max - 2)) / (1 - beta1 ** step)
            else:
                step_size = 1.0 / (1 - beta1 ** step)
            table, ctab, gc_rows = self._table(gi, plist)
            if gc_rows is not None:
                call("b21_grad_centralize", ptr(gc_rows), gc_rows.shape[0], stream_ptr())
            call("b21_ranger_step", ptr(table), ptr(ctab), ctab.shape[0], float(self.grad_scale), float(group["lr"]),
                 float(step_size), float(beta1), float(beta2), float(group["eps"]), float(group["weight_decay"]),
                 int(rect), int(step % group["k"] == 0), float(self.alpha), stream_ptr())
            # the kernel wrote the parameters through raw pointers: tell autograd / the packed-weight caches
            # (networks._B21Net._ensure_packed keys on ``_version``) that they changed in place
            torch._C._increment_version(plist)
        return loss

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._tables = {}

    def __setstate__(self, state):
        super().__setstate__(state)
        self._tables = {}
