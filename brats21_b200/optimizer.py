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
        self._dyn = self._dyn_buf = None

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

    def _plist(self, group):
        plist = [p for p in group["params"] if p.grad is not None]
        for p in plist:
            if not p.is_cuda:
                raise RuntimeError("brats21_b200.Ranger2020 runs on CUDA only (no CPU fallback)")
            if p.dtype != torch.float32 or p.grad.dtype != torch.float32 or not p.is_contiguous() \
                    or not p.grad.is_contiguous():
                raise RuntimeError("Ranger2020 fused step needs contiguous fp32 parameters and gradients")
        return plist

    def _advance(self, group, plist):
        """Host side of a step (optimizer.py:196-217): per-tensor step counters, N_sma, rectification, step size."""
        for p in plist:
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
        return step_size, bool(rect), step % group["k"] == 0

    # ---- CUDA-graph mode (engine.TrainStep): the step-dependent scalars live in device memory
    _RING = 16

    def enable_graph_mode(self):
        """After this, ``step()`` only LAUNCHES (its scalars come from a device buffer, so the launch can be captured
        and replayed) and ``begin_graph_step()`` — called eagerly before every replay — advances the host state and
        uploads the scalars.  Parameters must already have gradients and state (run one eager step first)."""
        if self._dyn_buf is None:
            dev = self.param_groups[0]["params"][0].device
            ng = len(self.param_groups)
            self._dyn_buf = torch.zeros((ng, 4), dtype=torch.float32, device=dev)
            self._dyn_host = torch.zeros((self._RING, ng, 4), dtype=torch.float32).pin_memory()
            self._dyn_events = [None] * self._RING
            self._dyn_slot = 0
        self._dyn = self._dyn_buf

    def disable_graph_mode(self):
        self._dyn = None

    def begin_graph_step(self):
        slot = self._dyn_slot
        self._dyn_slot = (slot + 1) % self._RING
        if self._dyn_events[slot] is not None:
            self._dyn_events[slot].synchronize()  # the upload that last used this pinned slot (RING steps ago)
        host = self._dyn_host[slot]
        for gi, group in enumerate(self.param_groups):
            plist = self._plist(group)
            if not plist:
                host[gi].zero_()
                continue
            step_size, rect, look = self._advance(group, plist)
            host[gi, 0], host[gi, 1], host[gi, 2], host[gi, 3] = step_size * group["lr"], float(rect), float(look), \
                float(self.grad_scale)
        self._dyn_buf.copy_(host, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self._dyn_events[slot] = ev

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        dyn = getattr(self, "_dyn", None)
        for gi, group in enumerate(self.param_groups):
            plist = self._plist(group)
            if not plist:
                continue
            beta1, beta2 = group["betas"]
            if dyn is None:
                step_size, rect, look = self._advance(group, plist)
            else:
                step_size, rect, look = 0.0, False, False  # overridden by dyn[gi] on the device
            table, ctab, gc_rows = self._table(gi, plist)
            if gc_rows is not None:
                call("b21_grad_centralize", ptr(gc_rows), gc_rows.shape[0], stream_ptr())
            call("b21_ranger_step", ptr(table), ptr(ctab), ctab.shape[0], float(self.grad_scale), float(group["lr"]),
                 float(step_size), float(beta1), float(beta2), float(group["eps"]), float(group["weight_decay"]),
                 int(rect), int(look), float(self.alpha), ptr(dyn[gi]) if dyn is not None else None, stream_ptr())
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
