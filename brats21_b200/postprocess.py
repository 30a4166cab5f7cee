"""Label-map post-processing on the GPU: the reference's CPU transforms right after the hot path.

  * ``KeepLargestConnectedComponent(threshold)``  utils/transforms.py:209-230 (+ get_largest_component :579-600)
  * ``ReplaceWithClosestValue(labels, thresh, axis)``  utils/transforms.py:233-268 (+ replace_w_closest_value_* :603-647)
  * ``ConvertToMultiChannelBasedOnBratsClasses``  MONAI, as chained in src/definer.py:679-692
  * ``pad_back``  shape_to_original + pad_back_to_shape_before_compose (utils/transforms.py:515-576)

Same constructor arguments and call convention as the reference transforms (label maps are [1, 1, D, H, W]); inputs
must be CUDA tensors — there is no CPU path here.
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import call, ptr, stream_ptr


def _label_u8(data: torch.Tensor) -> torch.Tensor:
    if not data.is_cuda:
        raise RuntimeError("postprocess: expected a CUDA tensor (no CPU fallback)")
    if data.dim() != 5 or data.shape[0] != 1 or data.shape[1] != 1:
        raise AssertionError("data shape must be 11HWD")
    return data.to(torch.uint8).contiguous() if data.dtype != torch.uint8 else data.contiguous().clone()


class KeepLargestConnectedComponent:
    """threshold=None keeps the largest component, otherwise the components with MORE than ``threshold`` voxels."""

    def __init__(self, threshold: Optional[int] = None) -> None:
        self.threshold = threshold

    def __call__(self, data: torch.Tensor) -> torch.Tensor:
        lab = _label_u8(data)
        d, h, w = lab.shape[2:]
        nbytes = _lib.load().b21_keep_components_workspace_bytes(d * h * w)
        work = torch.empty((nbytes,), dtype=torch.uint8, device=lab.device)
        call("b21_keep_components", ptr(lab), ptr(work), d, h, w, -1 if self.threshold is None else int(self.threshold),
             stream_ptr())
        return lab if data.dtype == torch.uint8 else lab.to(data.dtype)


class ReplaceWithClosestValue:
    """``labels`` is accepted and, as in the reference (utils/transforms.py:254-268), not used: every value with at
    most ``thresh`` voxels is replaced."""

    def __init__(self, labels: Sequence[int] = (3,), thresh: int = 20, axis: int = 2) -> None:
        self.labels, self.thresh, self.axis = labels, thresh, axis

    def __call__(self, data: torch.Tensor) -> torch.Tensor:
        lab = _label_u8(data)
        n0, n1, n2 = lab.shape[2:]
        nbytes = _lib.load().b21_replace_rare_workspace_bytes(int(self.thresh))
        work = torch.empty((nbytes,), dtype=torch.uint8, device=lab.device)
        call("b21_replace_rare_labels", ptr(lab), ptr(work), n0, n1, n2, int(self.thresh), int(self.axis), stream_ptr())
        return lab if data.dtype == torch.uint8 else lab.to(data.dtype)


def labels_to_channels(label: torch.Tensor) -> torch.Tensor:
    """BraTS label map [1, 1, D, H, W] (0/1/2/4) -> uint8 [1, 3, D, H, W] in (TC, WT, ET) order."""
    lab = _label_u8(label)
    d, h, w = lab.shape[2:]
    out = torch.empty((1, 3, d, h, w), dtype=torch.uint8, device=lab.device)
    call("b21_labels_to_channels", ptr(lab), ptr(out), d * h * w, stream_ptr())
    return out


def pad_back(x: torch.Tensor, meta) -> torch.Tensor:
    """Undo shape_to_divisible and CropForeground: [..., D', H', W'] -> [..., D, H, W] of the original image
    (zeros outside the foreground box).  Pure slicing/copy: device memory plumbing."""
    pb, pa = meta.pad_before, meta.pad_after
    sp = x.shape[-3:]
    core = x[..., pb[0]:sp[0] - pa[0], pb[1]:sp[1] - pa[1], pb[2]:sp[2] - pa[2]]
    out = torch.zeros(x.shape[:-3] + tuple(meta.original_shape), dtype=x.dtype, device=x.device)
    s, e = meta.start, meta.end
    out[..., s[0]:e[0], s[1]:e[1], s[2]:e[2]] = core
    return out
