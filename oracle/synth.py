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
from brats21_b200.synth import _gen, ellipsoid_mask, target, volume  # noqa: F401  (shared data generators)


def make_params(version: int, width: int = 48, seed: int = 123, perturb_affine: bool = True,
                inplanes: int = 4, num_classes: int = 3, norm: str = "group") -> Dict[str, torch.Tensor]:
    """Reference-format state_dict with the reference's init DISTRIBUTIONS (V1: kaiming-normal fan_out on convs,
    networks/factory.py:209-210; V2: torch defaults, equiunet2021.py:287) drawn from name-keyed generators.
    With perturb_affine the norm scales/offsets are jittered so that parity tests exercise them."""
    spec = nets.v1_param_shapes(width, inplanes, num_classes, norm) if version == 1 else \
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
        elif leaf == "running_mean":  # BatchNorm buffers: non-trivial values so that eval mode exercises them
            t = 0.2 * torch.randn(shape, generator=g)
        elif leaf == "running_var" and ".bn." in name:
            t = 0.5 + torch.rand(shape, generator=g)
        elif leaf == "num_batches_tracked":
            t = torch.zeros(shape, dtype=torch.long)
        elif leaf in ("v", "running_var"):
            t = torch.ones(shape)
        else:
            raise KeyError(name)
        out[name] = t if t.dtype == torch.long else t.float()
    return out
