"""Sliding-window inference with the reference's call signature (utils/inferers.py:26-40) on B200 kernels.

Window scheduling follows MONAI 0.6.0 ``dense_patch_slices`` / ``_get_scan_interval`` exactly (same window
origins, same order, same sw_batch_size grouping).  The importance map is the separable product of three 1-D
profiles (constant: ones; gaussian: erf-integrated Gaussian, sigma = sigma_scale * roi, truncated at 4 sigma,
normalised to 1 at the centre), accumulated in fp32 on the GPU by ``b21_blend_accumulate`` instead of on the CPU.

Fast path: when ``predictor`` is one of this package's networks the windows are cropped straight out of the
volume into channels-last bf16 (``b21_pack_windows``) and the deep-supervision heads are skipped (the reference
discards them at inferers.py:135-136).  Any other callable goes through the generic path (torch slicing for the
crop, the same CUDA blending).
"""
from __future__ import annotations

import collections
import itertools
import math
from typing import Any, Callable, Dict, List, Sequence, Tuple, Union

import torch
import torch.nn.functional as F

from . import ops

__all__ = ["sliding_window_inference"]


# ------------------------------------------------------------------------------------------ host-side geometry
def _fall_back_tuple(user, default):
    n = len(default)
    user = tuple(user) if isinstance(user, (list, tuple)) else (user,) * n
    if len(user) != n:
        raise ValueError(f"roi_size must have {n} elements")
    return tuple(u if (u and u > 0) else d for u, d in zip(user, default))


def _get_scan_interval(image_size: Sequence[int], roi_size: Sequence[int], num_spatial_dims: int,
                       overlap: float) -> Tuple[int, ...]:
    if len(image_size) != num_spatial_dims:
        raise ValueError("image coord different from spatial dims.")
    if len(roi_size) != num_spatial_dims:
        raise ValueError("roi coord different from spatial dims.")
    out = []
    for img, roi in zip(image_size, roi_size):
        if roi == img:
            out.append(int(roi))
        else:
            iv = int(roi * (1 - overlap))
            out.append(iv if iv > 0 else 1)
    return tuple(out)


def window_origins(image_size: Sequence[int], roi_size: Sequence[int], overlap: float) -> List[Tuple[int, ...]]:
    """Origins of all windows in MONAI order (first spatial dim slowest)."""
    roi = tuple(min(i, r) for i, r in zip(image_size, roi_size))
    interval = _get_scan_interval(image_size, roi, len(image_size), overlap)
    per_dim = []
    for img, r, iv in zip(image_size, roi, interval):
        num = int(math.ceil(float(img) / iv)) if iv else 1
        first = next((k for k in range(num) if k * iv + r >= img), None)
        count = first + 1 if first is not None else 1
        per_dim.append([k * iv - max(k * iv + r - img, 0) for k in range(count)])
    return list(itertools.product(*per_dim))


_profile_cache: Dict[Tuple, torch.Tensor] = {}


def importance_profiles(roi: Sequence[int], mode: str, sigma_scale, device) -> List[torch.Tensor]:
    """Per-axis 1-D factors of MONAI's importance map (fp32, on `device`)."""
    mode = getattr(mode, "value", mode)
    if mode not in ("constant", "gaussian"):
        raise ValueError(f"unsupported blend mode {mode!r}")
    ss = tuple(sigma_scale) if isinstance(sigma_scale, (list, tuple)) else (sigma_scale,) * len(roi)
    out = []
    for n, s in zip(roi, ss):
        key = (n, mode, float(s), str(device))
        if key not in _profile_cache:
            if mode == "constant":
                prof = torch.ones(n, dtype=torch.float64)
            else:
                sigma = n * s
                tail = int(max(sigma * 4.0, 0.5) + 0.5)
                x = torch.arange(n, dtype=torch.float64) - (n // 2)
                t = 0.70710678 / abs(sigma)
                prof = 0.5 * (torch.erf(t * (x + 0.5)) - torch.erf(t * (x - 0.5)))
                prof = torch.where(x.abs() <= tail, prof, torch.zeros_like(prof)).clamp(min=0)
                prof = prof / prof.max()
            _profile_cache[key] = prof.float().to(device)
        out.append(_profile_cache[key])
    return out


# One full-volume fp32 map per geometry: after CropForeground nearly every case has its own shape (and the axis
# permutations of the TTA add more), so the cache is a small LRU rather than a per-process leak.
_COUNT_CACHE_ENTRIES = 8
_count_cache: "collections.OrderedDict[Tuple, torch.Tensor]" = collections.OrderedDict()


_floor_cache: Dict[Tuple, float] = {}


def importance_floor(profiles) -> float:
    """MONAI clamps the 3-D importance map at its smallest non-zero value (compute_importance_map).  The map is the
    outer product of the profiles, so that value is the product of the per-axis smallest non-zero entries, and the clamp
    only changes entries where some factor is exactly 0 — possible only when the 4-sigma truncation falls inside the
    window (sigma_scale < 1/8).  Returns 0.0 (no clamp needed) when every profile is strictly positive."""
    key = tuple((p.data_ptr(), p.numel()) for p in profiles)  # the profiles are cached tensors: one host sync per geometry,
    if key not in _floor_cache:                                # not three per TTA variant (24 pipeline drains per volume)
        if all(bool((p > 0).all()) for p in profiles):
            _floor_cache[key] = 0.0
        else:
            mins = [p[p > 0].min() for p in profiles]
            _floor_cache[key] = float((mins[0] * mins[1]) * mins[2])
    return _floor_cache[key]


def count_map(image_size, roi, overlap, mode, sigma_scale, device) -> torch.Tensor:
    """Sum of importance weights over all windows ([1, D, H, W] fp32), cached per geometry."""
    mode = getattr(mode, "value", mode)
    key = (tuple(image_size), tuple(roi), float(overlap), mode, str(sigma_scale), str(device))
    if key not in _count_cache:
        cnt = torch.zeros((1,) + tuple(image_size), dtype=torch.float32, device=device)
        prof = importance_profiles(roi, mode, sigma_scale, device)
        origins = window_origins(image_size, roi, overlap)
        floor = importance_floor(prof)
        ops.blend_accumulate(None, cnt, prof, origins, floor)
        _count_cache[key] = cnt
        while len(_count_cache) > _COUNT_CACHE_ENTRIES:
            _count_cache.popitem(last=False)
    _count_cache.move_to_end(key)
    return _count_cache[key]


class WindowPlan:
    """Geometry of one sliding-window pass over an (augmented) image of spatial size `image_size0`."""

    def __init__(self, image_size0, roi_size, overlap, mode, sigma_scale, device):
        if overlap < 0 or overlap >= 1:
            raise AssertionError("overlap must be >= 0 and < 1.")
        self.image_size0 = tuple(int(s) for s in image_size0)
        roi = _fall_back_tuple(roi_size, self.image_size0)
        self.image_size = tuple(max(i, r) for i, r in zip(self.image_size0, roi))
        # F.pad amounts (inferers.py:101-109): half before, rest after
        self.pad_before = tuple((max(r - i, 0)) // 2 for i, r in zip(self.image_size0, roi))
        self.roi = tuple(min(i, r) for i, r in zip(self.image_size, roi))
        self.origins = window_origins(self.image_size, self.roi, overlap)
        self.profiles = importance_profiles(self.roi, mode, sigma_scale, device)
        self.wfloor = importance_floor(self.profiles)
        self.count = count_map(self.image_size, self.roi, overlap, mode, sigma_scale, device)


def _is_b21_net(predictor) -> bool:
    return hasattr(predictor, "forward_packed") and hasattr(predictor, "pack_input")


def accumulate_windows(vol: torch.Tensor, vol_idx: int, net, plan: WindowPlan, sw_batch_size: int, acc: torch.Tensor,
                       variant=((0, 1, 2), (0, 0, 0))):
    """Fast path: run `net` on every window of the augmented view `variant` of vol[vol_idx] and blend into `acc`
    ([K, *plan.image_size] fp32, zeroed by the caller)."""
    perm, flip = variant
    d, h, w = plan.roi
    ws = net._ws.setdefault(("sw", sw_batch_size, d, h, w), {})
    for g0 in range(0, len(plan.origins), sw_batch_size):
        group = plan.origins[g0:g0 + sw_batch_size]
        nb = len(group)
        x8 = net._buf(ws, f"win{nb}", (nb, d, h, w, 8))
        shifted = [tuple(o - p for o, p in zip(org, plan.pad_before)) for org in group]
        ops.pack_windows(vol, x8, shifted, perm=perm, flip=flip, vol_index=[vol_idx] * nb)
        logits = net.forward_infer(x8)
        ops.blend_accumulate(logits, acc, plan.profiles, group, plan.wfloor)


def sliding_window_inference(
        inputs: torch.Tensor,
        roi_size: Union[Sequence[int], int],
        sw_batch_size: int,
        predictor: Callable[..., torch.Tensor],
        overlap: float = 0.25,
        mode: str = "constant",
        sigma_scale: Union[Sequence[float], float] = 0.125,
        padding_mode: str = "constant",
        cval: float = 0.0,
        sw_device: Union[torch.device, str, None] = None,
        device: Union[torch.device, str, None] = None,
        *args: Any,
        **kwargs: Any,
) -> torch.Tensor:
    """Drop-in for utils/inferers.py:26 (5-D inputs [N, C, D, H, W]); returns fp32 [N, K, D, H, W] on `device`."""
    if inputs.dim() != 5:
        raise ValueError("this implementation handles 3-D volumes: inputs must be [N, C, D, H, W]")
    if overlap < 0 or overlap >= 1:
        raise AssertionError("overlap must be >= 0 and < 1.")
    if not inputs.is_cuda:
        raise RuntimeError("sliding_window_inference runs on CUDA only (no CPU fallback)")
    out_device = inputs.device if device is None else torch.device(device)
    padding_mode = getattr(padding_mode, "value", padding_mode)
    mode = getattr(mode, "value", mode)
    nb = inputs.shape[0]
    plan = WindowPlan(inputs.shape[2:], roi_size, overlap, mode, sigma_scale, inputs.device)
    fast = _is_b21_net(predictor) and padding_mode == "constant" and cval == 0.0 and not args and not kwargs \
        and not predictor.training
    acc = None
    if fast:
        vol = inputs.detach().to(torch.float32).contiguous()
        with torch.no_grad():
            predictor._ensure_packed()
            acc = torch.zeros((nb, predictor.num_classes) + plan.image_size, dtype=torch.float32, device=inputs.device)
            if nb == 1:
                accumulate_windows(vol, 0, predictor, plan, sw_batch_size, acc[0])
            else:
                _fast_multi_volume(vol, predictor, plan, sw_batch_size, acc)
    else:
        pad = []
        for k in range(4, 1, -1):
            diff = max(plan.image_size[k - 2] - inputs.shape[k], 0)
            half = diff // 2
            pad.extend([half, diff - half])
        x = F.pad(inputs, pad=pad, mode=padding_mode, value=cval) if any(pad) else inputs
        total = len(plan.origins) * nb
        r = plan.roi
        for g0 in range(0, total, sw_batch_size):
            idxs = list(range(g0, min(g0 + sw_batch_size, total)))
            wins = []
            for idx in idxs:
                b, o = idx // len(plan.origins), plan.origins[idx % len(plan.origins)]
                wins.append(x[b:b + 1, :, o[0]:o[0] + r[0], o[1]:o[1] + r[1], o[2]:o[2] + r[2]])
            seg = predictor(torch.cat(wins), *args, **kwargs)
            while isinstance(seg, (tuple, list)):  # deep supervision: first output only
                seg = seg[0]
            seg = seg.detach().to(device=inputs.device, dtype=torch.float32).contiguous()
            if acc is None:
                acc = torch.zeros((nb, seg.shape[1]) + plan.image_size, dtype=torch.float32, device=inputs.device)
            for j, idx in enumerate(idxs):
                b, o = idx // len(plan.origins), plan.origins[idx % len(plan.origins)]
                ops.blend_accumulate(seg[j:j + 1], acc[b], plan.profiles, [o], plan.wfloor)
    out = torch.empty((nb, acc.shape[1]) + plan.image_size0, dtype=torch.float32, device=inputs.device)
    for b in range(nb):
        ops.tta_accumulate(acc[b], plan.count, out[b], pad_before=plan.pad_before, apply_sigmoid=False, overwrite=True)
    return out.to(out_device)


def _fast_multi_volume(vol, net, plan, sw_batch_size, acc):
    """Batched volumes: windows are grouped across volumes exactly as the reference does (inferers.py:126-131)."""
    nwin = len(plan.origins)
    total = nwin * vol.shape[0]
    d, h, w = plan.roi
    ws = net._ws.setdefault(("sw", sw_batch_size, d, h, w), {})
    for g0 in range(0, total, sw_batch_size):
        idxs = list(range(g0, min(g0 + sw_batch_size, total)))
        group = [plan.origins[i % nwin] for i in idxs]
        vidx = [i // nwin for i in idxs]
        x8 = net._buf(ws, f"win{len(idxs)}", (len(idxs), d, h, w, 8))
        shifted = [tuple(o - p for o, p in zip(org, plan.pad_before)) for org in group]
        ops.pack_windows(vol, x8, shifted, vol_index=vidx)
        logits = net.forward_infer(x8)
        for j, (o, b) in enumerate(zip(group, vidx)):
            ops.blend_accumulate(logits[j:j + 1], acc[b], plan.profiles, [o], plan.wfloor)
