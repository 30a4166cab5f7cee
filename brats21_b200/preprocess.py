"""Input side of the hot path on the GPU.

``crop_normalize_pad`` = the reference's test-time data transforms between NIfTI decode and the network
(src/definer.py:561-567 ``CropForegroundd`` + ``NormalizeIntensityd(nonzero=True, channel_wise=True)``, optionally
``remove_outliers`` as in training, definer.py:466-467) followed by ``shape_to_divisible(k=8)``
(utils/transforms.py:483-512, called at learning/engine.py:229): one bounding-box reduction, one statistics
reduction and one fused normalise + crop + pad pass, all on the device.  The only host round trip is the six
bounding-box integers, which fix the output shape (the reference's shapes are data dependent as well).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Tuple

import torch

from ._lib import call, ptr, stream_ptr


@dataclass
class CropMeta:
    """What ``pad_back_to_shape_before_compose`` / ``shape_to_original`` need (utils/transforms.py:515-576)."""
    original_shape: Tuple[int, int, int]
    start: Tuple[int, int, int]   # foreground_start_coord
    end: Tuple[int, int, int]     # foreground_end_coord
    pad_before: Tuple[int, int, int]
    pad_after: Tuple[int, int, int]


def foreground_bbox(img: torch.Tensor) -> torch.Tensor:
    """Device int32[6] = (start d, h, w, end d, h, w) of ``img > 0`` over any channel (MONAI CropForeground)."""
    _check(img)
    c, d, h, w = img.shape
    bbox = torch.empty((6,), dtype=torch.int32, device=img.device)
    call("b21_foreground_bbox", ptr(img), c, d, h, w, ptr(bbox), stream_ptr())
    return bbox


def nonzero_stats(img: torch.Tensor, bbox: torch.Tensor) -> torch.Tensor:
    """Device float64 [C, 3] = (count, sum, sum of squares) of the non-zero voxels inside ``bbox``."""
    _check(img)
    c, d, h, w = img.shape
    stats = torch.empty((c, 3), dtype=torch.float64, device=img.device)
    call("b21_nonzero_stats", ptr(img), c, d, h, w, ptr(bbox), ptr(stats), stream_ptr())
    return stats


def crop_normalize_pad(img: torch.Tensor, k: int = 8, remove_outliers: bool = False, outliers_value: float = 3.0,
                       crop_foreground: bool = True):
    """img: raw intensities [C, D, H, W] fp32 on CUDA.  Returns (volume [1, C, D', H', W'] fp32 with D', H', W'
    multiples of ``k``, CropMeta)."""
    _check(img)
    c, d, h, w = img.shape
    if crop_foreground:
        bbox = foreground_bbox(img)
        b = bbox.tolist()  # the one host sync: the output shape depends on the data
        if b[3] <= b[0] or b[4] <= b[1] or b[5] <= b[2]:  # no foreground: keep the whole volume
            b = [0, 0, 0, d, h, w]
            bbox = torch.tensor(b, dtype=torch.int32, device=img.device)
    else:
        b = [0, 0, 0, d, h, w]
        bbox = torch.tensor(b, dtype=torch.int32, device=img.device)
    stats = nonzero_stats(img, bbox)
    box = [b[3] - b[0], b[4] - b[1], b[5] - b[2]]
    pad = [(-s) % k for s in box]
    pb = [(p + 1) // 2 for p in pad]
    pa = [p // 2 for p in pad]
    od, oh, ow = (s + p for s, p in zip(box, pad))
    out = torch.empty((1, c, od, oh, ow), dtype=torch.float32, device=img.device)
    call("b21_normalize_crop_pad", ptr(img), ptr(out), c, d, h, w, ptr(bbox), ptr(stats), od, oh, ow, pb[0], pb[1],
         pb[2], float(outliers_value) if remove_outliers else 0.0, stream_ptr())
    return out, CropMeta((d, h, w), tuple(b[:3]), tuple(b[3:]), tuple(pb), tuple(pa))


def _check(img):
    if not (img.is_cuda and img.dtype == torch.float32 and img.dim() == 4 and img.is_contiguous()):
        raise RuntimeError("preprocess: expected a contiguous CUDA fp32 [C, D, H, W] volume (no CPU fallback)")
    if img.shape[0] > 16:
        raise RuntimeError("preprocess: at most 16 channels")
