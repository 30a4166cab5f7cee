"""Test-time augmentation with the reference's ttach-style protocol (tta/base.py, tta/transforms.py):
``Compose([...])`` is an iterable of ``Transformer`` objects exposing ``augment_image`` / ``deaugment_mask`` /
``deaugment_label`` and ``len()``; parameters are combined as a cartesian product, the image chain runs the
transforms in order and the mask chain runs their inverses in reverse order (tta/base.py:112-131).

Every geometric transform here is a signed axis permutation, so each ``Transformer`` also carries ``variant`` =
(perm, flip) — the description the CUDA kernels (b21_pack_windows / b21_tta_accumulate) use to read windows of the
augmented volume straight out of the source volume and to scatter predictions back, without ever materialising an
augmented copy.  The tensor methods remain available (views only) for drop-in use with arbitrary predictors.
"""
from __future__ import annotations

import itertools
from typing import Callable, List, Optional, Sequence, Tuple

import torch


class BaseTransform:
    identity_param = None

    def __init__(self, name: str, params: Sequence):
        self.pname = name
        self.params = params

    def apply_aug_image(self, image, **kw):
        raise NotImplementedError

    def apply_deaug_mask(self, mask, **kw):
        raise NotImplementedError

    def apply_deaug_label(self, label, **kw):
        return label


class DualTransform(BaseTransform):
    pass


class ImageOnlyTransform(BaseTransform):
    def apply_deaug_mask(self, mask, **kw):
        return mask


class OnAxes(DualTransform):
    """Re-order the spatial axes: "zxy" (identity), "xyz", "yzx" (tta/transforms.py:16-49)."""
    identity_param = "zxy"
    _FWD = {"zxy": None, "xyz": (0, 1, 3, 4, 2), "yzx": (0, 1, 4, 2, 3)}
    _INV = {"zxy": None, "xyz": (0, 1, 4, 2, 3), "yzx": (0, 1, 3, 4, 2)}

    def __init__(self, axes: List[str]):
        assert all(a in self._FWD for a in axes), "axes need to be 'xyz', 'yzx', 'zxy'"
        super().__init__("axe", axes)

    def apply_aug_image(self, image, axe="zxy", **kw):
        return image if self._FWD[axe] is None else image.permute(*self._FWD[axe])

    def apply_deaug_mask(self, mask, axe="zxy", **kw):
        return mask if self._INV[axe] is None else mask.permute(*self._INV[axe])


class _Flip(DualTransform):
    identity_param = False
    dim = 3

    def __init__(self):
        super().__init__("apply", [False, True])

    def apply_aug_image(self, image, apply=False, **kw):
        return image.flip(self.dim) if apply else image

    def apply_deaug_mask(self, mask, apply=False, **kw):
        return mask.flip(self.dim) if apply else mask


class HorizontalFlip(_Flip):
    """flip(3) (tta/transforms.py:52-73)."""
    dim = 3


class VerticalFlip(_Flip):
    """flip(2) (tta/transforms.py:76-98)."""
    dim = 2


class DepthFlip(_Flip):
    """flip(4): not in the reference library; completes the 8 axis-flip set BASELINE.json names."""
    dim = 4


class Rotate90(DualTransform):
    """rot90 by 0/90/180/270 degrees in the (2, 3) plane (tta/transforms.py:149-173)."""
    identity_param = 0

    def __init__(self, angles: List[int]):
        if self.identity_param not in angles:
            angles = [self.identity_param] + list(angles)
        super().__init__("angle", angles)

    @staticmethod
    def _k(angle):
        return angle // 90 if angle >= 0 else (angle + 360) // 90

    def apply_aug_image(self, image, angle=0, **kw):
        return torch.rot90(image, self._k(angle), (2, 3))

    def apply_deaug_mask(self, mask, angle=0, **kw):
        return torch.rot90(mask, self._k(-angle), (2, 3))


class Transformer:
    """One TTA variant.  ``variant`` = (perm, flip): augmented[a] = volume[s], s_j = flip_j ? dim_j-1-a_perm_j :
    a_perm_j — None when the chain is not a pure signed permutation (then only the tensor methods can be used)."""

    def __init__(self, aug_chain: List[Callable], mask_chain: List[Callable], label_chain: List[Callable]):
        self._aug, self._mask, self._label = aug_chain, mask_chain, label_chain
        self.variant: Optional[Tuple[Tuple[int, int, int], Tuple[int, int, int]]] = _probe_variant(self)

    def augment_image(self, image):
        for f in self._aug:
            image = f(image)
        return image

    def deaugment_mask(self, mask):
        for f in self._mask:
            mask = f(mask)
        return mask

    def deaugment_label(self, label):
        for f in self._label:
            label = f(label)
        return label


def _probe_variant(tr: "Transformer"):
    """Derive (perm, flip) by pushing coordinate grids of an asymmetric 2x3x4 volume through the image chain."""
    dims = (2, 3, 4)
    grids = torch.meshgrid(*[torch.arange(s) for s in dims], indexing="ij")
    probe = torch.stack(grids).unsqueeze(0).float()  # [1, 3, 2, 3, 4]; channel j holds source coordinate j
    try:
        aug = tr.augment_image(probe)
    except Exception:  # noqa: BLE001 - non-geometric transform
        return None
    if aug.dim() != 5 or sorted(aug.shape[2:]) != sorted(dims):
        return None
    perm, flip = [], []
    for j in range(3):
        cj = aug[0, j]
        found = None
        for ax in range(3):
            if aug.shape[2 + ax] != dims[j]:
                continue
            idx = torch.arange(dims[j]).float().reshape([-1 if k == ax else 1 for k in range(3)]).expand_as(cj)
            if torch.equal(cj, idx):
                found = (ax, 0)
            elif torch.equal(cj, dims[j] - 1 - idx):
                found = (ax, 1)
        if found is None:
            return None
        perm.append(found[0])
        flip.append(found[1])
    if sorted(perm) != [0, 1, 2]:
        return None
    # the mask chain must be the exact inverse
    if not torch.equal(tr.deaugment_mask(aug), probe):
        return None
    return tuple(perm), tuple(flip)


class Compose:
    def __init__(self, transforms: List[BaseTransform]):
        self.aug_transforms = transforms
        self.aug_transform_parameters = list(itertools.product(*[t.params for t in transforms]))
        self.deaug_transforms = transforms[::-1]
        self.deaug_transform_parameters = [p[::-1] for p in self.aug_transform_parameters]

    def __iter__(self):
        for aug_p, deaug_p in zip(self.aug_transform_parameters, self.deaug_transform_parameters):
            aug = [_bind(t.apply_aug_image, t.pname, p) for t, p in zip(self.aug_transforms, aug_p)]
            mask = [_bind(t.apply_deaug_mask, t.pname, p) for t, p in zip(self.deaug_transforms, deaug_p)]
            label = [_bind(t.apply_deaug_label, t.pname, p) for t, p in zip(self.deaug_transforms, deaug_p)]
            yield Transformer(aug, mask, label)

    def __len__(self):
        return len(self.aug_transform_parameters)


def _bind(fn, name, value):
    def bound(x):
        return fn(x, **{name: value})
    return bound


def get_tta_transforms() -> Compose:
    """The reference's inference TTA (src/definer.py:647-658): 2 axis orders x h-flip x 4 rotations = 16."""
    return Compose([OnAxes(axes=["zxy", "xyz"]), HorizontalFlip(), Rotate90(angles=[0, 90, 180, 270])])


def get_flip8_transforms() -> Compose:
    """All 8 subsets of axis flips (BASELINE.json config 3: '8-flip TTA')."""
    return Compose([VerticalFlip(), HorizontalFlip(), DepthFlip()])
