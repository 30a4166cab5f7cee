"""Deterministic synthetic inputs: there is no network access for the BraTS data, so every arm
(kernels, oracle, bench) sees these seeded volumes.  Nothing in the reference fixes them; they are fixed here."""
from __future__ import annotations

import zlib
from typing import Sequence

import torch


def _gen(seed: int, name: str) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((zlib.crc32(name.encode()) ^ (seed * 2654435761)) & 0x7FFFFFFFFFFFFFFF)
    return g


def ellipsoid_mask(shape: Sequence[int], semi=(0.9, 0.8, 0.85), center_shift=(0.0, 0.0, 0.0)) -> torch.Tensor:
    axes = []
    for n, s, c in zip(shape, semi, center_shift):
        half = (n - 1) / 2.0
        axes.append(((torch.arange(n, dtype=torch.float32) - half - c * half) / (s * half + 1e-6)) ** 2)
    r2 = axes[0].reshape(-1, 1, 1) + axes[1].reshape(1, -1, 1) + axes[2].reshape(1, 1, -1)
    return (r2 <= 1.0)


def volume(seed: int = 0, shape=(240, 240, 155), channels: int = 4) -> torch.Tensor:
    """[1, C, D, H, W] fp32: z-scored-looking noise clipped to [-3, 3] inside an ellipsoid 'brain', exact zeros
    outside (exercises remove_background_voxels, utils/transforms.py:536-550)."""
    g = _gen(seed, "volume")
    x = torch.randn((1, channels) + tuple(shape), generator=g).clamp_(-3.0, 3.0)
    # low-frequency structure so that the network sees something other than white noise
    coarse = torch.randn((1, channels) + tuple(max(2, s // 16) for s in shape), generator=g)
    x = 0.6 * x + torch.nn.functional.interpolate(coarse, size=tuple(shape), mode="trilinear", align_corners=True)
    x = x.clamp_(-3.0, 3.0)
    x[x == 0] = 1e-3
    return x * ellipsoid_mask(shape).to(x.dtype)[None, None]


def target(shape=(128, 128, 128)) -> torch.Tensor:
    """[1, 3, D, H, W] fp32 {0,1}: nested ellipsoids ET ⊂ TC ⊂ WT in MONAI channel order (TC, WT, ET)."""
    wt = ellipsoid_mask(shape, (0.55, 0.5, 0.6), (0.1, -0.1, 0.05))
    tc = ellipsoid_mask(shape, (0.35, 0.3, 0.4), (0.1, -0.1, 0.05))
    et = ellipsoid_mask(shape, (0.2, 0.15, 0.25), (0.1, -0.1, 0.05))
    return torch.stack([tc, wt, et]).float()[None]
