"""TEST INFRASTRUCTURE ONLY — CPU/fp32 oracle for the inference wrappers around the networks.

Restates, in plain torch/numpy:
  * sliding_window_inference + _get_scan_interval         utils/inferers.py:26-186
      (with MONAI 0.6.0's dense_patch_slices / compute_importance_map, SURVEY.md Appendix A — parity unpinned)
  * the ttach-style TTA algebra (OnAxes / HorizontalFlip / VerticalFlip / Rotate90 and Compose's product and
    reversed de-augmentation order)                        tta/base.py:103-136, tta/transforms.py:16-173
  * Engine._apply_tta + the sigmoid / mean / >=0.5 ensemble  learning/engine.py:424-440,236-249
  * label post-processing: ConvertToBratsClassesBasedOnMultiChannel, ChangeLabel3To4, remove_background_voxels,
    shape_to_divisible / shape_to_original                 utils/transforms.py:169-206,483-550
"""
from __future__ import annotations

import itertools
import math
from typing import Callable, List, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F


# ------------------------------------------------------------------------------------------ sliding window
def scan_interval(image_size: Sequence[int], roi_size: Sequence[int], overlap: float) -> Tuple[int, ...]:
    """utils/inferers.py:165-186."""
    if len(image_size) != len(roi_size):
        raise ValueError("image coord different from spatial dims.")
    out = []
    for img, roi in zip(image_size, roi_size):
        if roi == img:
            out.append(int(roi))
        else:
            iv = int(roi * (1 - overlap))
            out.append(iv if iv > 0 else 1)
    return tuple(out)


def window_starts(image_size: Sequence[int], roi_size: Sequence[int], interval: Sequence[int]) -> List[List[int]]:
    """Per-dimension window start lists of MONAI dense_patch_slices."""
    starts = []
    for img, roi, iv in zip(image_size, roi_size, interval):
        if iv == 0:
            count = 1
        else:
            num = int(math.ceil(float(img) / iv))
            first = next((d for d in range(num) if d * iv + roi >= img), None)
            count = first + 1 if first is not None else 1
        dim = []
        for idx in range(count):
            s = idx * iv
            s -= max(s + roi - img, 0)
            dim.append(s)
        starts.append(dim)
    return starts


def window_grid(image_size, roi_size, overlap) -> List[Tuple[int, ...]]:
    """All window origins, first spatial dim slowest (meshgrid 'ij' flattened row-major)."""
    roi = tuple(min(i, r) for i, r in zip(image_size, roi_size))
    starts = window_starts(image_size, roi, scan_interval(image_size, roi, overlap))
    return list(itertools.product(*starts))


def gaussian_profile(n: int, sigma_scale: float = 0.125) -> torch.Tensor:
    """1-D factor of MONAI's gaussian importance map for a window of n voxels: erf-integrated Gaussian of
    sigma = n*sigma_scale centred on voxel n//2, truncated at 4 sigma, normalised to 1 at the centre."""
    sigma = n * sigma_scale
    tail = int(max(sigma * 4.0, 0.5) + 0.5)
    x = torch.arange(n, dtype=torch.float64) - (n // 2)
    t = 0.70710678 / abs(sigma)
    k = 0.5 * (torch.erf(t * (x + 0.5)) - torch.erf(t * (x - 0.5)))
    k = torch.where(x.abs() <= tail, k, torch.zeros_like(k)).clamp(min=0)
    return (k / k.max())


def importance_map(roi: Sequence[int], mode: str = "constant", sigma_scale: float = 0.125) -> torch.Tensor:
    if mode == "constant":
        return torch.ones(tuple(roi), dtype=torch.float32)
    if mode != "gaussian":
        raise ValueError(f"unknown blend mode {mode}")
    prof = [gaussian_profile(n, sigma_scale) for n in roi]
    m = prof[0].reshape(-1, 1, 1) * prof[1].reshape(1, -1, 1) * prof[2].reshape(1, 1, -1)
    m = (m / m.max()).float()
    nz = m[m != 0].min()
    return torch.clamp(m, min=float(nz))


def sliding_window_inference(inputs: torch.Tensor, roi_size, sw_batch_size: int, predictor: Callable,
                             overlap: float = 0.25, mode: str = "constant", sigma_scale: float = 0.125,
                             cval: float = 0.0) -> torch.Tensor:
    """utils/inferers.py:26-162 for 5-D inputs with constant padding; fp32 accumulation."""
    if overlap < 0 or overlap >= 1:
        raise AssertionError("overlap must be >= 0 and < 1.")
    nb = inputs.shape[0]
    img0 = list(inputs.shape[2:])
    roi = tuple(r if (r and r > 0) else i for r, i in zip(_rep(roi_size, 3), img0))
    image_size = tuple(max(i, r) for i, r in zip(img0, roi))
    pad = []
    for k in range(4, 1, -1):
        diff = max(roi[k - 2] - inputs.shape[k], 0)
        half = diff // 2
        pad.extend([half, diff - half])
    x = F.pad(inputs, pad=pad, mode="constant", value=cval)
    origins = window_grid(image_size, roi, overlap)
    roi_v = tuple(min(i, r) for i, r in zip(image_size, roi))
    wmap = importance_map(roi_v, mode, sigma_scale).to(inputs.device)
    total = len(origins) * nb
    out = cnt = None
    for g0 in range(0, total, sw_batch_size):
        idxs = range(g0, min(g0 + sw_batch_size, total))
        wins = []
        for idx in idxs:
            b, o = idx // len(origins), origins[idx % len(origins)]
            wins.append(x[b:b + 1, :, o[0]:o[0] + roi_v[0], o[1]:o[1] + roi_v[1], o[2]:o[2] + roi_v[2]])
        prob = predictor(torch.cat(wins))
        while isinstance(prob, (tuple, list)):  # deep supervision: first output only (inferers.py:135-136)
            prob = prob[0]
        prob = prob.float()
        if out is None:
            out = torch.zeros((nb, prob.shape[1]) + image_size, dtype=torch.float32, device=prob.device)
            cnt = torch.zeros_like(out)
        for j, idx in enumerate(idxs):
            b, o = idx // len(origins), origins[idx % len(origins)]
            sl = (b, slice(None), slice(o[0], o[0] + roi_v[0]), slice(o[1], o[1] + roi_v[1]),
                  slice(o[2], o[2] + roi_v[2]))
            out[sl] += wmap * prob[j]
            cnt[sl] += wmap
    out = out / cnt
    return out[:, :, pad[4]:pad[4] + img0[0], pad[2]:pad[2] + img0[1], pad[0]:pad[0] + img0[2]]


def _rep(v, n):
    return tuple(v) if isinstance(v, (list, tuple)) else (v,) * n


# ------------------------------------------------------------------------------------------ TTA
class Variant:
    """One TTA variant as an (augment, de-augment) pair of tensor functions on [N,C,D,H,W]."""

    def __init__(self, name, aug, deaug):
        self.name, self.augment_image, self.deaugment_mask = name, aug, deaug


def _chain(fs):
    def run(x):
        for f in fs:
            x = f(x)
        return x
    return run


def _on_axes(axe):  # tta/transforms.py:32-46
    if axe == "zxy":
        return (lambda x: x), (lambda x: x)
    if axe == "xyz":
        return (lambda x: x.permute(0, 1, 3, 4, 2)), (lambda x: x.permute(0, 1, 4, 2, 3))
    if axe == "yzx":
        return (lambda x: x.permute(0, 1, 4, 2, 3)), (lambda x: x.permute(0, 1, 3, 4, 2))
    raise AssertionError("axes need to be 'xyz', 'yzx', 'zxy'")


def _flip(dim, apply):
    f = (lambda x: x.flip(dim)) if apply else (lambda x: x)
    return f, f


def _rot90(angle):  # tta/transforms.py:165-170
    def k_of(a):
        return a // 90 if a >= 0 else (a + 360) // 90
    return (lambda x: torch.rot90(x, k_of(angle), (2, 3))), (lambda x: torch.rot90(x, k_of(-angle), (2, 3)))


def compose(transforms: Sequence[Tuple[str, Sequence]]) -> List[Variant]:
    """tta.Compose: cartesian product of the per-transform parameter lists; the image chain applies the
    transforms in order, the mask chain applies their inverses in reverse order (tta/base.py:112-131).
    `transforms` is a list of (kind, params) with kind in {"axes","hflip","vflip","rot90","flip"}; "flip" takes
    (dim, bool) parameters and exists for the north-star's 8 axis-flip variants."""
    makers = {"axes": _on_axes, "hflip": lambda a: _flip(3, a), "vflip": lambda a: _flip(2, a), "rot90": _rot90,
              "flip": lambda p: _flip(p[0], p[1])}
    out = []
    for combo in itertools.product(*[params for _, params in transforms]):
        pairs = [makers[kind](p) for (kind, _), p in zip(transforms, combo)]
        aug = _chain([a for a, _ in pairs])
        deaug = _chain([d for _, d in reversed(pairs)])
        out.append(Variant("|".join(f"{k}={p}" for (k, _), p in zip(transforms, combo)), aug, deaug))
    return out


def reference_tta() -> List[Variant]:
    """src/definer.py:647-658 — 2 axes x 2 hflip x 4 rotations = 16 variants."""
    return compose([("axes", ["zxy", "xyz"]), ("hflip", [False, True]), ("rot90", [0, 90, 180, 270])])


def flip8_tta() -> List[Variant]:
    """BASELINE.json's '8-flip TTA': every subset of the three spatial axes flipped."""
    return compose([("flip", [(2, False), (2, True)]), ("flip", [(3, False), (3, True)]),
                    ("flip", [(4, False), (4, True)])])


def apply_tta(forward: Callable, img: torch.Tensor, variants: Sequence[Variant]) -> List[torch.Tensor]:
    """Engine._apply_tta (learning/engine.py:424-440), first head only."""
    outs = []
    for v in variants:
        y = forward(v.augment_image(img))
        while isinstance(y, (tuple, list)):
            y = y[0]
        outs.append(v.deaugment_mask(y))
    return outs


def ensemble_mean_threshold(logit_list: Sequence[torch.Tensor], thresh: float = 0.5):
    """sigmoid each, mean over variants x models, threshold (engine.py:239-249; AsDiscrete(threshold_values))."""
    prob = torch.stack([torch.sigmoid(l.float()) for l in logit_list]).mean(dim=0)
    return prob, (prob >= thresh).float()


# ------------------------------------------------------------------------------------------ post-processing
def remove_background_voxels(img: torch.Tensor, outputs: torch.Tensor) -> torch.Tensor:
    """utils/transforms.py:536-550: zero predictions where all input channels are exactly 0."""
    mask = (img != 0).any(dim=1, keepdim=False).to(outputs.dtype)
    return outputs * mask.unsqueeze(1)


def brats_label_map(onehot: torch.Tensor, et_label: int = 4) -> torch.Tensor:
    """ConvertToBratsClassesBasedOnMultiChannel + ChangeLabel3To4 (utils/transforms.py:169-206).
    Channels are (TC, WT, ET); result uint8 [1,1,D,H,W] with NCR/NET=1, ED=2, ET=4."""
    assert onehot.dim() == 5 and onehot.shape[0] == 1 and onehot.shape[1] == 3
    tc, wt, et = onehot[0, 0].bool(), onehot[0, 1].bool(), onehot[0, 2].bool()
    lab = torch.zeros(tc.shape, dtype=torch.uint8, device=onehot.device)
    lab[et] = 3
    lab[tc & ~et] = 1
    lab[wt & ~tc] = 2
    lab[lab == 3] = et_label
    return lab[None, None]


def shape_to_divisible(data: torch.Tensor, k: int = 8):
    """utils/transforms.py:483-512: zero-pad spatial dims up to a multiple of k (ceil half before)."""
    shp = np.array(data.shape[-3:])
    tgt = np.ceil(shp / k).astype(int) * k
    p = tgt - shp
    pb, pa = np.ceil(p / 2).astype(int), np.floor(p / 2).astype(int)
    out = F.pad(data, (int(pb[2]), int(pa[2]), int(pb[1]), int(pa[1]), int(pb[0]), int(pa[0])))
    return out, pb, pa


def shape_to_original(data: torch.Tensor, pb, pa) -> torch.Tensor:
    shp = np.array(data.shape[-3:])
    up = shp - pa
    return data[..., pb[0]:up[0], pb[1]:up[1], pb[2]:up[2]].contiguous()
