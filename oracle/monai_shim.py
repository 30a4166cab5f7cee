"""TEST INFRASTRUCTURE ONLY — never imported by the product path.

Minimal stand-in for the parts of `monai==0.6.0` (requirements.txt:17 of the reference; NOT vendored under
/root/reference and not installable here) that the reference hot path imports, so that the UNMODIFIED reference
modules (networks/*, utils/inferers.py, tta/*, learning/optimizer.py) can be imported in this container to
generate golden vectors (tests/golden/make_golden.py) and to cross-check the oracle port.

Semantics restated from MONAI 0.6.0 as documented in SURVEY.md Appendix A.  PARITY UNPINNED: MONAI's own source
and tests are not available offline, so these restatements cannot be diffed against the original.
"""
from __future__ import annotations

import math
import sys
import types
from enum import Enum

import torch
import torch.nn as nn
import torch.nn.functional as F


# ----------------------------------------------------------------------------- monai.utils
class BlendMode(Enum):
    CONSTANT = "constant"
    GAUSSIAN = "gaussian"


class PytorchPadMode(Enum):
    CONSTANT = "constant"
    REFLECT = "reflect"
    REPLICATE = "replicate"
    CIRCULAR = "circular"


def ensure_tuple_rep(v, n):
    if isinstance(v, (list, tuple)):
        if len(v) != n:
            raise ValueError(f"sequence must have length {n}, got {len(v)}")
        return tuple(v)
    return (v,) * n


def fall_back_tuple(user_provided, default, func=lambda x: x and x > 0):
    ndim = len(default)
    user = ensure_tuple_rep(user_provided, ndim)
    return tuple(u if func(u) else d for u, d in zip(user, default))


# ----------------------------------------------------------------------------- monai.data.utils
def get_valid_patch_size(image_size, patch_size):
    ndim = len(image_size)
    patch_size_ = ensure_tuple_rep(patch_size, ndim)
    return tuple(min(ms, ps or ms) for ms, ps in zip(image_size, patch_size_))


def dense_patch_slices(image_size, patch_size, scan_interval):
    num_spatial_dims = len(image_size)
    patch_size = get_valid_patch_size(image_size, patch_size)
    scan_interval = ensure_tuple_rep(scan_interval, num_spatial_dims)
    scan_num = []
    for i in range(num_spatial_dims):
        if scan_interval[i] == 0:
            scan_num.append(1)
        else:
            num = int(math.ceil(float(image_size[i]) / scan_interval[i]))
            scan_dim = next((d for d in range(num) if d * scan_interval[i] + patch_size[i] >= image_size[i]), None)
            scan_num.append(scan_dim + 1 if scan_dim is not None else 1)
    starts = []
    for dim in range(num_spatial_dims):
        dim_starts = []
        for idx in range(scan_num[dim]):
            start_idx = idx * scan_interval[dim]
            start_idx -= max(start_idx + patch_size[dim] - image_size[dim], 0)
            dim_starts.append(start_idx)
        starts.append(dim_starts)
    grids = torch.meshgrid(*[torch.tensor(s) for s in starts], indexing="ij")
    out = torch.stack([g.reshape(-1) for g in grids], dim=1).tolist()
    return [tuple(slice(s, s + patch_size[d]) for d, s in enumerate(x)) for x in out]


def _gaussian_1d_erf(sigma: float, truncated: float = 4.0) -> torch.Tensor:
    tail = int(max(float(sigma) * truncated, 0.5) + 0.5)
    x = torch.arange(-tail, tail + 1, dtype=torch.float)
    t = 0.70710678 / abs(float(sigma))
    out = 0.5 * ((t * (x + 0.5)).erf() - (t * (x - 0.5)).erf())
    out = out.clamp(min=0)
    return out / out.sum()


def _separable_gaussian(x: torch.Tensor, sigmas) -> torch.Tensor:
    """monai.networks.layers.GaussianFilter(approx='erf'): per-axis zero-padded 'same' convolution."""
    nd = x.dim() - 2
    for ax in range(nd):
        k = _gaussian_1d_erf(sigmas[ax]).to(x)
        shape = [1, 1] + [1] * nd
        shape[2 + ax] = -1
        pad = [0] * nd
        pad[ax] = (k.numel() - 1) // 2
        conv = [F.conv1d, F.conv2d, F.conv3d][nd - 1]
        x = conv(x, k.reshape(shape), padding=pad)
    return x


def compute_importance_map(patch_size, mode=BlendMode.CONSTANT, sigma_scale=0.125, device="cpu"):
    mode = BlendMode(mode)
    device = torch.device(device)
    if mode == BlendMode.CONSTANT:
        return torch.ones(patch_size, device=device).float()
    center = [i // 2 for i in patch_size]
    sigma_scale = ensure_tuple_rep(sigma_scale, len(patch_size))
    sigmas = [i * s for i, s in zip(patch_size, sigma_scale)]
    imap = torch.zeros(patch_size, device=device)
    imap[tuple(center)] = 1
    imap = _separable_gaussian(imap[None, None], sigmas)[0, 0]
    imap = imap / torch.max(imap)
    imap = imap.float()
    min_non_zero = imap[imap != 0].min().item()
    return torch.clamp(imap, min=min_non_zero)


# ----------------------------------------------------------------------------- monai.networks
def same_padding(kernel_size, dilation=1):
    if isinstance(kernel_size, (list, tuple)):
        raise NotImplementedError("shim: scalar kernel sizes only")
    p = (kernel_size - 1) / 2.0 * dilation
    if p != int(p):
        raise NotImplementedError("same padding not available for this k/d")
    return int(p)


class _Factory:
    def __init__(self, table):
        self._t = table

    def __getitem__(self, key):
        if isinstance(key, tuple):
            name, dim = key
            return self._t[str(name).upper()][dim]
        return self._t[str(key).upper()]


class _ConvFactory(_Factory):
    CONV = "CONV"
    CONVTRANS = "CONVTRANS"


Conv = _ConvFactory({"CONV": {1: nn.Conv1d, 2: nn.Conv2d, 3: nn.Conv3d},
                     "CONVTRANS": {1: nn.ConvTranspose1d, 2: nn.ConvTranspose2d, 3: nn.ConvTranspose3d}})
Act = _Factory({"RELU": nn.ReLU, "LEAKYRELU": nn.LeakyReLU, "ELU": nn.ELU, "PRELU": nn.PReLU,
                "SIGMOID": nn.Sigmoid, "TANH": nn.Tanh})


class MaxAvgPool(nn.Module):
    def __init__(self, spatial_dims, kernel_size, stride=None, padding=0, ceil_mode=False):
        super().__init__()
        assert spatial_dims == 3
        self.max_pool = nn.MaxPool3d(kernel_size, stride, padding, ceil_mode=ceil_mode)
        self.avg_pool = nn.AvgPool3d(kernel_size, stride, padding, ceil_mode=ceil_mode)

    def forward(self, x):
        return torch.cat([self.max_pool(x), self.avg_pool(x)], dim=1)


def _act(spec):
    if isinstance(spec, (tuple, list)):
        name, kw = spec
        return Act[name](**kw)
    return Act[spec]()


class ChannelSELayer(nn.Module):
    def __init__(self, spatial_dims, in_channels, r=2, acti_type_1=("relu", {"inplace": True}),
                 acti_type_2="sigmoid"):
        super().__init__()
        assert spatial_dims == 3
        self.avg_pool = nn.AdaptiveAvgPool3d(1)
        channels = int(in_channels // r)
        self.fc = nn.Sequential(nn.Linear(in_channels, channels, bias=True), _act(acti_type_1),
                                nn.Linear(channels, in_channels, bias=True), _act(acti_type_2))

    def forward(self, x):
        b, c = x.shape[:2]
        y = self.avg_pool(x).view(b, c)
        y = self.fc(y).view([b, c] + [1] * (x.dim() - 2))
        return x * y


class ResidualSELayer(ChannelSELayer):
    def forward(self, x):
        return x + super().forward(x)


class Randomizable:
    R = None

    def set_random_state(self, seed=None, state=None):
        import numpy as np
        self.R = np.random.RandomState(seed)
        return self

    def randomize(self, data=None):
        raise NotImplementedError


import contextlib


class DiceLoss(nn.Module):
    """monai.losses.DiceLoss restated for the flag combinations the reference uses (src/definer.py:184-212; SURVEY.md
    Appendix A): include_background=True, to_onehot_y=False, sigmoid / no activation, squared_pred, jaccard, batch,
    reduction mean|sum|none."""

    def __init__(self, include_background=True, to_onehot_y=False, sigmoid=False, softmax=False, other_act=None,
                 squared_pred=False, jaccard=False, reduction="mean", smooth_nr=1e-5, smooth_dr=1e-5, batch=False):
        super().__init__()
        if to_onehot_y or softmax or other_act is not None or not include_background:
            raise NotImplementedError("monai shim: DiceLoss flag combination not restated")
        self.sigmoid, self.squared_pred, self.jaccard, self.batch = sigmoid, squared_pred, jaccard, batch
        self.reduction = getattr(reduction, "value", reduction)
        self.smooth_nr, self.smooth_dr = float(smooth_nr), float(smooth_dr)

    def forward(self, input, target):  # noqa: A002
        if self.sigmoid:
            input = torch.sigmoid(input)  # noqa: A001
        if target.shape != input.shape:
            raise AssertionError(f"ground truth has differing shape ({target.shape}) from input ({input.shape})")
        reduce_axis = list(range(2, input.dim()))
        if self.batch:
            reduce_axis = [0] + reduce_axis
        intersection = torch.sum(target * input, dim=reduce_axis)
        if self.squared_pred:
            target = torch.pow(target, 2)
            input = torch.pow(input, 2)  # noqa: A001
        denominator = torch.sum(target, dim=reduce_axis) + torch.sum(input, dim=reduce_axis)
        if self.jaccard:
            denominator = 2.0 * (denominator - intersection)
        f = 1.0 - (2.0 * intersection + self.smooth_nr) / (denominator + self.smooth_dr)
        if self.reduction == "mean":
            return torch.mean(f)
        if self.reduction == "sum":
            return torch.sum(f)
        return f


@contextlib.contextmanager
def eval_mode(*nets):
    """monai.networks.utils.eval_mode (learning/engine.py:18,226): no_grad + .eval(), training flags restored."""
    training = [n for n in nets if n.training]
    try:
        with torch.no_grad():
            yield [n.eval() for n in nets]
    finally:
        for n in training:
            n.train()


def install():
    """Inject the shim as `monai` into sys.modules (idempotent)."""
    if "monai" in sys.modules and getattr(sys.modules["monai"], "__b21_shim__", False):
        return

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    monai = mod("monai", __b21_shim__=True, __version__="0.6.0-shim")
    monai.utils = mod("monai.utils", BlendMode=BlendMode, PytorchPadMode=PytorchPadMode,
                      fall_back_tuple=fall_back_tuple, ensure_tuple_rep=ensure_tuple_rep)
    monai.data = mod("monai.data")
    monai.data.utils = mod("monai.data.utils", compute_importance_map=compute_importance_map,
                           dense_patch_slices=dense_patch_slices, get_valid_patch_size=get_valid_patch_size)
    monai.networks = mod("monai.networks")
    monai.networks.blocks = mod("monai.networks.blocks", MaxAvgPool=MaxAvgPool, ResidualSELayer=ResidualSELayer,
                                ChannelSELayer=ChannelSELayer)
    monai.networks.layers = mod("monai.networks.layers", same_padding=same_padding, Act=Act, Conv=Conv)
    monai.networks.layers.factories = mod("monai.networks.layers.factories", Act=Act, Conv=Conv)
    monai.networks.utils = mod("monai.networks.utils", eval_mode=eval_mode)
    monai.losses = mod("monai.losses", DiceLoss=DiceLoss)
    monai.losses.dice = mod("monai.losses.dice", DiceLoss=DiceLoss)
    monai.transforms = mod("monai.transforms")
    monai.transforms.compose = mod("monai.transforms.compose", Randomizable=Randomizable)
