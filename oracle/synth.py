"""TEST INFRASTRUCTURE ONLY — deterministic synthetic inputs and weights shared by the oracle, the golden
generator, the parity tests and bench.py.  Nothing in the reference fixes these; they are
fixed here so that every arm sees identical data.
"""
from __future__ import annotations

import math
import zlib
from typing import Dict, Sequence

import torch

from . import nets


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


def make_params(version: int, width: int = 48, seed: int = 123, perturb_affine: bool = True,
                inplanes: int = 4, num_classes: int = 3) -> Dict[str, torch.Tensor]:
    """Reference-format state_dict with the reference's init DISTRIBUTIONS (V1: kaiming-normal fan_out on convs,
    networks/factory.py:209-210; V2: torch defaults, equiunet2021.py:287) drawn from name-keyed generators.
    With perturb_affine the norm scales/offsets are jittered so that parity tests exercise them."""
    spec = nets.v1_param_shapes(width, inplanes, num_classes) if version == 1 else \
        nets.v2_param_shapes(width, inplanes, num_classes)
    out: Dict[str, torch.Tensor] = {}
    for name, shape in spec:
        g = _gen(seed, name)
        leaf = name.rsplit(".", 1)[-1]
        if leaf == "weight" and len(shape) >= 2:  # conv / linear
            fan_in = shape[1] * int(math.prod(shape[2:]))
            fan_out = shape[0] * int(math.prod(shape[2:]))
            if version == 1 and len(shape) == 5:
                t = torch.randn(shape, generator=g) * math.sqrt(2.0 / fan_out)
            else:
                b = 1.0 / math.sqrt(fan_in)
                t = (torch.rand(shape, generator=g) * 2 - 1) * b
        elif leaf == "bias" and ".bn." not in name:
            wshape = dict(spec)[name[:-4] + "weight"]
            fan_in = wshape[1] * int(math.prod(wshape[2:]))
            b = 1.0 / math.sqrt(fan_in)
            t = (torch.rand(shape, generator=g) * 2 - 1) * b
        elif leaf in ("gamma",) or (leaf == "weight" and ".bn." in name):
            t = torch.ones(shape)
            if perturb_affine:
                t = t + 0.1 * torch.randn(shape, generator=g)
        elif leaf in ("beta",) or (leaf == "bias" and ".bn." in name):
            t = torch.zeros(shape)
            if perturb_affine:
                t = 0.1 * torch.randn(shape, generator=g)
        elif leaf in ("v", "running_var"):
            t = torch.ones(shape)
        else:
            raise KeyError(name)
        out[name] = t.float()
    return out
