"""TEST INFRASTRUCTURE ONLY — CPU/fp32 oracle for the two networks on the hot path.

Functional restatement (plain torch fp32 ops over a reference-format ``state_dict``) of
  * V1  EquiUnet           networks/equiunet2020.py:408-500  (UBlock/ConvBnRelu :51-123, GroupNorm(8) factory.py:182)
  * V2  EquiUnetASSPEvo    networks/equiunet2021.py:225-333  (EvoNorm3D-S0 :48-105, ASPP :121-189,
                                                              ConvEvoBlockCorrected/ConvEvo :192-222)
plus the MONAI 0.6.0 blocks they use (MaxAvgPool, ResidualSELayer).

Pinned against the unmodified reference modules through tests/golden/*.npz (produced by tests/golden/make_golden.py
from the unmodified reference; checked by tests/test_oracle_golden.py everywhere).  MONAI-derived pieces remain "parity unpinned" (source not vendored).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch
import torch.nn.functional as F

Params = Dict[str, torch.Tensor]


# ------------------------------------------------------------------------------------------ building blocks
def group_norm_relu(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, groups: int = 8,
                    eps: float = 1e-5) -> torch.Tensor:
    """GroupNorm(8, C, affine) + ReLU  (factory.py:182, equiunet2020.py:60-61)."""
    n, c = x.shape[:2]
    xg = x.reshape(n, groups, -1)
    mean = xg.mean(dim=2, keepdim=True)
    var = xg.var(dim=2, unbiased=False, keepdim=True)
    y = ((xg - mean) * torch.rsqrt(var + eps)).reshape(x.shape)
    y = y * gamma.reshape(1, c, 1, 1, 1) + beta.reshape(1, c, 1, 1, 1)
    return torch.relu(y)


def evonorm_s0(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, groups: int = 8,
               eps: float = 1e-5) -> torch.Tensor:
    """EvoNorm3D S0, efficient swish variant: x*sigmoid(x) / sqrt(var_unbiased_group(x)+eps) * gamma + beta
    (equiunet2021.py:48-52,95-105; `v` and `running_var` never enter the result)."""
    n, c = x.shape[:2]
    xg = x.reshape(n, groups, -1)
    var = xg.var(dim=2, unbiased=True, keepdim=True)
    std = torch.sqrt(var + eps)  # [n, groups, 1]
    num = (x * torch.sigmoid(x)).reshape(n, groups, -1)
    y = (num / std).reshape(x.shape)
    return y * gamma.reshape(1, c, 1, 1, 1) + beta.reshape(1, c, 1, 1, 1)


def residual_se(x: torch.Tensor, w1, b1, w2, b2) -> torch.Tensor:
    """MONAI ResidualSELayer(r=2, relu, sigmoid): x + x * sigmoid(W2 relu(W1 mean(x) + b1) + b2)."""
    m = x.mean(dim=(2, 3, 4))
    s = torch.sigmoid(F.linear(torch.relu(F.linear(m, w1, b1)), w2, b2))
    return x + x * s.reshape(*s.shape, 1, 1, 1)


def max_avg_pool(x: torch.Tensor) -> torch.Tensor:
    """MONAI MaxAvgPool(kernel 2): channel concat of max-pool and avg-pool."""
    return torch.cat([F.max_pool3d(x, 2), F.avg_pool3d(x, 2)], dim=1)


def up_trilinear(x: torch.Tensor, scale: int = 2) -> torch.Tensor:
    return F.interpolate(x, scale_factor=scale, mode="trilinear", align_corners=True)


# ------------------------------------------------------------------------------------------ V1
def activation(y: torch.Tensor, act: str) -> torch.Tensor:
    """get_act (factory.py:195-200) through MONAI's Act table: relu | leakyrelu (slope 0.01) | elu (alpha 1)."""
    if act == "relu":
        return torch.relu(y)
    if act == "leakyrelu":
        return F.leaky_relu(y, 0.01)
    if act == "elu":
        return F.elu(y)
    raise KeyError(act)


def _v1_cbr(p: Params, key: str, x: torch.Tensor, dil: int = 1, norm: str = "group", act: str = "relu",
            training: bool = False) -> torch.Tensor:
    """ConvBnRelu (equiunet2020.py:51-75) for every norm of get_norm_layer (factory.py:179-192) except "bcn"."""
    if norm == "none":
        return activation(F.conv3d(x, p[f"{key}.conv.weight"], p[f"{key}.conv.bias"], padding=dil, dilation=dil), act)
    y = F.conv3d(x, p[f"{key}.conv.weight"], None, padding=dil, dilation=dil)
    g, b = p[f"{key}.bn.weight"], p[f"{key}.bn.bias"]
    if norm == "group":
        if act == "relu":
            return group_norm_relu(y, g, b)
        return activation(F.group_norm(y, 8, g, b, 1e-5), act)
    if norm == "instance":  # nn.InstanceNorm3d(affine=True): per (n, c) statistics, biased variance, eps 1e-5
        return activation(F.instance_norm(y, None, None, g, b, True, 0.1, 1e-5), act)
    if norm == "batch":     # nn.BatchNorm3d(affine=True); training updates the running statistics in place
        rm, rv = p[f"{key}.bn.running_mean"], p[f"{key}.bn.running_var"]
        return activation(F.batch_norm(y, rm, rv, g, b, training, 0.1, 1e-5), act)
    raise ValueError("Norm type is not correct")


def _v1_ublock(p: Params, key: str, x: torch.Tensor, dils=(1, 1), **kw) -> torch.Tensor:
    x = _v1_cbr(p, f"{key}.ConvBnRelu1", x, dils[0], **kw)
    return _v1_cbr(p, f"{key}.ConvBnRelu2", x, dils[1], **kw)


def _head(p: Params, key: str, x: torch.Tensor, scale: int) -> torch.Tensor:
    y = F.conv3d(x, p[f"{key}.weight"], p[f"{key}.bias"])
    return up_trilinear(y, scale) if scale > 1 else y


def equiunet_v1_forward(p: Params, x: torch.Tensor, deep_supervision: bool = True, norm: str = "group",
                        act: str = "relu", training: bool = False):
    """equiunet2020.py:467-500."""
    kw = dict(norm=norm, act=act, training=training)
    down1 = _v1_ublock(p, "encoder1", x, **kw)
    down2 = _v1_ublock(p, "encoder2", F.max_pool3d(down1, 2), **kw)
    down3 = _v1_ublock(p, "encoder3", F.max_pool3d(down2, 2), **kw)
    down4 = _v1_ublock(p, "encoder4", F.max_pool3d(down3, 2), **kw)
    bottom = _v1_ublock(p, "bottom", down4, (2, 2), **kw)
    bottom_2 = _v1_cbr(p, "bottom_2", torch.cat([down4, bottom], dim=1), **kw)
    up3 = _v1_ublock(p, "decoder3", torch.cat([down3, up_trilinear(bottom_2)], dim=1), **kw)
    up2 = _v1_ublock(p, "decoder2", torch.cat([down2, up_trilinear(up3)], dim=1), **kw)
    up1 = _v1_ublock(p, "decoder1", torch.cat([down1, up_trilinear(up2)], dim=1), **kw)
    out = _head(p, "outconv", up1, 1)
    if not deep_supervision:
        return out
    deeps = [_head(p, "deep_bottom.0", bottom, 8), _head(p, "deep_bottom2.0", bottom_2, 8),
             _head(p, "deep3.0", up3, 4), _head(p, "deep2.0", up2, 2)]
    return out, deeps


# ------------------------------------------------------------------------------------------ V2
def _v2_block(p: Params, key: str, x: torch.Tensor) -> torch.Tensor:
    k = f"{key}.conv_conv_se"
    x = F.conv3d(x, p[f"{k}.0.weight"], p[f"{k}.0.bias"], padding=1)
    x = evonorm_s0(x, p[f"{k}.1.gamma"], p[f"{k}.1.beta"])
    x = F.conv3d(x, p[f"{k}.3.weight"], p[f"{k}.3.bias"], padding=1)
    x = evonorm_s0(x, p[f"{k}.4.gamma"], p[f"{k}.4.beta"])
    return residual_se(x, p[f"{k}.6.fc.0.weight"], p[f"{k}.6.fc.0.bias"], p[f"{k}.6.fc.2.weight"],
                       p[f"{k}.6.fc.2.bias"])


def _v2_convevo(p: Params, key: str, x: torch.Tensor) -> torch.Tensor:
    y = F.conv3d(x, p[f"{key}.conv.weight"], p[f"{key}.conv.bias"])
    return evonorm_s0(y, p[f"{key}.evo.gamma"], p[f"{key}.evo.beta"])


def _v2_aspp(p: Params, x: torch.Tensor) -> torch.Tensor:
    outs = []
    for i, dil in enumerate((1, 2, 4, 6)):
        w, b = p[f"aspp.convs.{i}.weight"], p[f"aspp.convs.{i}.bias"]
        pad = 0 if w.shape[2] == 1 else dil
        outs.append(F.conv3d(x, w, b, padding=pad, dilation=dil))
    return _v2_convevo(p, "aspp.conv_k1", torch.cat(outs, dim=1))


def equiunet_v2_forward(p: Params, x: torch.Tensor, deep_supervision: bool = True):
    """equiunet2021.py:289-333."""
    down1 = _v2_block(p, "encoder1", x)
    down2 = _v2_block(p, "encoder2", max_avg_pool(down1))
    down3 = _v2_block(p, "encoder3", max_avg_pool(down2))
    down4 = _v2_block(p, "encoder4", max_avg_pool(down3))
    assp = _v2_aspp(p, down4)
    down1b = _v2_convevo(p, "bridge1", down1)
    down2b = _v2_convevo(p, "bridge2", down2)
    down3b = _v2_convevo(p, "bridge3", down3)
    up3 = _v2_block(p, "decoder3", torch.cat([down3b, up_trilinear(_v2_convevo(p, "upconv3", assp))], dim=1))
    up2 = _v2_block(p, "decoder2", torch.cat([down2b, up_trilinear(_v2_convevo(p, "upconv2", up3))], dim=1))
    up1 = _v2_block(p, "decoder1", torch.cat([down1b, up_trilinear(_v2_convevo(p, "upconv1", up2))], dim=1))
    out = _head(p, "out_conv", up1, 1)
    if not deep_supervision:
        return out
    return out, [_head(p, "deep3.0", up3, 4), _head(p, "deep2.0", up2, 2)]


# ------------------------------------------------------------------------------------------ parameter specs
def v1_param_shapes(width: int = 48, inplanes: int = 4, num_classes: int = 3,
                    norm: str = "group") -> List[Tuple[str, Tuple[int, ...]]]:
    """state_dict entries of the reference EquiUnet in registration order; `norm` as
    get_norm_layer (factory.py:179-192): "none" gives the convs a bias and no `bn`, "batch" adds the BatchNorm buffers."""
    f = [width * 2 ** i for i in range(4)]
    out: List[Tuple[str, Tuple[int, ...]]] = []

    def cbr(key, cin, cout):
        out.append((f"{key}.conv.weight", (cout, cin, 3, 3, 3)))
        if norm == "none":
            out.append((f"{key}.conv.bias", (cout,)))
            return
        out.append((f"{key}.bn.weight", (cout,)))
        out.append((f"{key}.bn.bias", (cout,)))
        if norm == "batch":
            out.append((f"{key}.bn.running_mean", (cout,)))
            out.append((f"{key}.bn.running_var", (cout,)))
            out.append((f"{key}.bn.num_batches_tracked", ()))

    def ublock(key, cin, mid, cout):
        cbr(f"{key}.ConvBnRelu1", cin, mid)
        cbr(f"{key}.ConvBnRelu2", mid, cout)

    ublock("encoder1", inplanes, f[0], f[0])
    ublock("encoder2", f[0], f[1], f[1])
    ublock("encoder3", f[1], f[2], f[2])
    ublock("encoder4", f[2], f[3], f[3])
    ublock("bottom", f[3], f[3], f[3])
    cbr("bottom_2", f[3] * 2, f[2])
    ublock("decoder3", f[2] * 2, f[2], f[1])
    ublock("decoder2", f[1] * 2, f[1], f[0])
    ublock("decoder1", f[0] * 2, f[0], f[0])
    for key, cin in (("outconv", f[0]), ("deep_bottom.0", f[3]), ("deep_bottom2.0", f[2]), ("deep3.0", f[1]),
                     ("deep2.0", f[0])):
        out.append((f"{key}.weight", (num_classes, cin, 1, 1, 1)))
        out.append((f"{key}.bias", (num_classes,)))
    return out


def v2_param_shapes(width: int = 48, inplanes: int = 4, num_classes: int = 3) -> List[Tuple[str, Tuple[int, ...]]]:
    """state_dict entries (parameters AND the running_var buffers) of the reference EquiUnetASSPEvo."""
    f = [width * 2 ** i for i in range(4)]
    out: List[Tuple[str, Tuple[int, ...]]] = []

    def conv(key, cin, cout, k):
        out.append((f"{key}.weight", (cout, cin, k, k, k)))
        out.append((f"{key}.bias", (cout,)))

    def evo(key, c):
        for name in ("gamma", "beta", "v", "running_var"):
            out.append((f"{key}.{name}", (1, c, 1, 1, 1)))

    def block(key, cin, cout):
        k = f"{key}.conv_conv_se"
        conv(f"{k}.0", cin, cout, 3)
        evo(f"{k}.1", cout)
        conv(f"{k}.3", cout, cout, 3)
        evo(f"{k}.4", cout)
        out.append((f"{k}.6.fc.0.weight", (cout // 2, cout)))
        out.append((f"{k}.6.fc.0.bias", (cout // 2,)))
        out.append((f"{k}.6.fc.2.weight", (cout, cout // 2)))
        out.append((f"{k}.6.fc.2.bias", (cout,)))

    def convevo(key, cin, cout):
        conv(f"{key}.conv", cin, cout, 1)
        evo(f"{key}.evo", cout)

    block("encoder1", inplanes, f[0])
    block("encoder2", 2 * f[0], f[1])
    block("encoder3", 2 * f[1], f[2])
    block("encoder4", 2 * f[2], f[3])
    convevo("bridge1", f[0], f[0] // 2)
    convevo("bridge2", f[1], f[1] // 2)
    convevo("bridge3", f[2], f[2] // 2)
    for i, k in enumerate((1, 3, 3, 3)):
        conv(f"aspp.convs.{i}", f[3], f[3] // 4, k)
    convevo("aspp.conv_k1", f[3], f[3])
    convevo("upconv3", f[3], f[3] // 4)
    block("decoder3", f[2], f[2])
    convevo("upconv2", f[2], f[2] // 4)
    block("decoder2", f[1], f[1])
    convevo("upconv1", f[1], f[1] // 4)
    block("decoder1", f[0], f[0])
    conv("out_conv", f[0], num_classes, 1)
    conv("deep3.0", f[2], num_classes, 1)
    conv("deep2.0", f[1], num_classes, 1)
    return out
