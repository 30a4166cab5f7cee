"""Drop-in EquiUnet (V1) and EquiUnetASSPEvo (V2) whose forward runs on the sm_100a kernels of libb21.so.

Constructor signatures, ``forward`` return structure and ``state_dict`` keys/shapes are those of the reference
(networks/equiunet2020.py:408-500, networks/equiunet2021.py:225-333), so reference
checkpoints load unchanged.  torch.nn modules are used ONLY as parameter holders (same construction order as the
reference, hence the same random init under the same seed); their forward is never called — there is no eager
fallback, and a missing CUDA library raises.

Internally activations are channels-last bf16 ([N, D, H, W, C]); channel concatenations are written in place as
channel slices of one buffer; norm statistics come out of the conv epilogue.
"""
from __future__ import annotations

import warnings
from collections import OrderedDict
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn

from . import ops


# ================================================================================================ holders
def _conv3(cin, cout, dil=1, bias=False):
    return nn.Conv3d(cin, cout, kernel_size=3, padding=dil, dilation=dil, bias=bias)


def _conv1(cin, cout, bias=True):
    return nn.Conv3d(cin, cout, kernel_size=1, bias=bias)


_NORMS = {"group": lambda c: nn.GroupNorm(8, c, affine=True), "instance": lambda c: nn.InstanceNorm3d(c, affine=True),
          "batch": lambda c: nn.BatchNorm3d(c, affine=True), "none": None}   # get_norm_layer, factory.py:179-192
_ACTS = {"relu": lambda: nn.ReLU(inplace=True), "leakyrelu": lambda: nn.LeakyReLU(inplace=True),
         "elu": lambda: nn.ELU(inplace=True)}                                 # get_act, factory.py:195-200 (MONAI Act)


class ConvBnRelu(nn.Sequential):
    """Parameter holder mirroring equiunet2020.py:51-75: conv (bias only without a norm) -> norm -> act -> dropout,
    with the reference's sub-module names (the activation is registered under its own name)."""

    def __init__(self, cin, cout, act, dil=1, dropout=0.0, norm="group"):
        mods = [("conv", _conv3(cin, cout, dil, bias=_NORMS[norm] is None))]
        if _NORMS[norm] is not None:
            mods.append(("bn", _NORMS[norm](cout)))
        mods += [(act, _ACTS[act]()), ("dropout", nn.Dropout(p=dropout))]
        super().__init__(OrderedDict(mods))
        self.dil = dil


class UBlock(nn.Sequential):
    def __init__(self, cin, mid, cout, act, dils=(1, 1), dropout=0.0, norm="group"):
        super().__init__(OrderedDict([("ConvBnRelu1", ConvBnRelu(cin, mid, act, dils[0], dropout, norm)),
                                      ("ConvBnRelu2", ConvBnRelu(mid, cout, act, dils[1], dropout, norm))]))


class EvoNorm3D(nn.Module):
    """Parameter holder for EvoNorm3D-S0 (equiunet2021.py:55-105): gamma, beta, (unused) v and running_var."""

    def __init__(self, c):
        super().__init__()
        self.gamma = nn.Parameter(torch.ones(1, c, 1, 1, 1))
        self.beta = nn.Parameter(torch.zeros(1, c, 1, 1, 1))
        self.v = nn.Parameter(torch.ones(1, c, 1, 1, 1))
        self.register_buffer("running_var", torch.ones(1, c, 1, 1, 1))
        self.eps = 1e-5


class ResidualSELayer(nn.Module):
    """Parameter holder for MONAI ResidualSELayer(r=2): keys fc.0.*, fc.2.*."""

    def __init__(self, c, r=2):
        super().__init__()
        self.fc = nn.Sequential(nn.Linear(c, c // r), nn.ReLU(inplace=True), nn.Linear(c // r, c), nn.Sigmoid())


class ConvEvoBlockCorrected(nn.Module):
    def __init__(self, cin, cout, dropout):
        super().__init__()
        self.conv_conv_se = nn.Sequential(_conv3(cin, cout, bias=True), EvoNorm3D(cout), nn.Dropout(dropout),
                                          _conv3(cout, cout, bias=True), EvoNorm3D(cout), nn.Dropout(dropout),
                                          ResidualSELayer(cout))


class ConvEvo(nn.Module):
    def __init__(self, cin, cout, dropout):
        super().__init__()
        self.conv = _conv1(cin, cout)
        self.evo = EvoNorm3D(cout)
        self.drop = nn.Dropout(dropout)


class SimpleASPPEVO(nn.Module):
    def __init__(self, cin, cbranch, kernel_sizes=(1, 3, 3, 3), dilations=(1, 2, 4, 6)):
        super().__init__()
        if len(kernel_sizes) != len(dilations):
            raise ValueError("kernel_sizes and dilations length must match, "
                             f"got kernel_sizes={len(kernel_sizes)} dilations={len(dilations)}.")
        self.dilations = tuple(dilations)
        self.convs = nn.ModuleList()
        for k, d in zip(kernel_sizes, dilations):
            self.convs.append(nn.Conv3d(cin, cbranch, kernel_size=k, dilation=d, padding=(k - 1) // 2 * d))
        self.conv_k1 = ConvEvo(cbranch * len(kernel_sizes), cbranch * len(kernel_sizes), 0)


def _head(cin, ncls, scale):
    return nn.Sequential(_conv1(cin, ncls), nn.Upsample(scale_factor=scale, mode="trilinear", align_corners=True))


# ================================================================================================ runtime base
class _B21Net(nn.Module):
    """Shared runtime: weight packing cache, workspace cache, input packing."""

    def _init_runtime(self):
        self._packed: Dict[str, object] = {}
        self._aliases: List[Tuple[torch.Tensor, torch.Tensor]] = []  # (fp32 view or copy, source parameter)
        self._pack_key = None
        self._ws: Dict[Tuple, Dict[str, torch.Tensor]] = {}
        self.skip_deep_heads_in_eval = False  # set by the inference wrappers (they discard the deep heads)
        self._gs = None  # flat gradient store of the training path (brats21_b200.autograd.GradStore)
        self._graphs: Dict[Tuple, Tuple] = {}  # CUDA graphs of the inference forward, keyed by (input buffer, shape)
        self._pack_table = None  # (number of packed convs, device job table, jobs, blocks) of _refresh_packed

    # ---- training path (brats21_b200/autograd.py)
    def grad_store(self):
        if self._gs is None:
            from .autograd import GradStore
            self._gs = GradStore(self, self._grad_order())
        return self._gs

    def _grad_order(self):
        raise NotImplementedError(f"{type(self).__name__}: training is not on the accelerated path yet")

    def _forward_train(self, x8, want_deep):
        raise NotImplementedError(f"{type(self).__name__}: training is not on the accelerated path yet")

    def _backward_train(self, tape, dout, ddeeps, gs):
        raise NotImplementedError(f"{type(self).__name__}: training is not on the accelerated path yet")

    # ---- packing
    def _param_key(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    def _ensure_packed(self):
        """Packed bf16 conv weights follow the parameters: rebuilt when a parameter's storage changed, re-packed IN
        PLACE (same buffers, so captured CUDA graphs stay valid) when only its version counter moved — an optimizer
        step (Ranger2020.step bumps the counters after its fused raw-pointer update), ``load_state_dict`` / ``copy_``."""
        key = self._param_key()
        if key == self._pack_key:
            return
        if self._pack_key is not None and len(key) == len(self._pack_key) and \
                all(a[0] == b[0] for a, b in zip(key, self._pack_key)):
            self._refresh_packed()
        else:
            self._packed, self._aliases, self._graphs, self._pack_table = {}, [], {}, None
            self._pack()
            self._sync_pack_table()
        self._pack_key = key

    def _sync_pack_table(self):
        """(packed convs, device job table of b21_pack_batch), rebuilt whenever packed convs were added (host -> device
        copy: call it outside CUDA-graph capture — _ensure_packed and the training-path packers do)."""
        items = [v for v in self._packed.values() if isinstance(v, ops.PackedConv)]
        table = self._pack_table
        if table is None or table[0] != len(items):
            groups: Dict[int, list] = {}  # images made from the same fp32 source (forward + transposed packings)
            for v in items:
                for j in v.pack_jobs():
                    groups.setdefault(j.w, []).append(j)
            jobs, blk = [], 0
            for members in groups.values():  # one block per 16 x 16 (cout, cin) source tile, shared by the group
                nblk = ((members[0].cout + 15) // 16) * ((members[0].cin + 15) // 16)
                for j in members:
                    j.blk0, j.nblk = blk, nblk
                jobs += members
                blk += nblk
            raw = (ops._lib.PackJob * len(jobs))(*jobs)
            host = torch.frombuffer(bytearray(bytes(raw)), dtype=torch.uint8)
            table = self._pack_table = (len(items), host.to(self._device()), len(jobs), blk)
        return items, table

    def _refresh_packed(self):
        """Re-pack every conv weight IN PLACE with one launch (csrc/pack_batch.cu): ~90 packing launches of 2-30 us took
        0.41-0.56 ms at the head of every training step (measured as separate launches and as parallel graph branches)."""
        items, table = self._sync_pack_table()
        for v in items:
            v.sync_sources()
        if table[2]:
            ops.call("b21_pack_batch", ops.ptr(table[1]), table[2], table[3], ops.stream_ptr())
        for v in items:
            if "ws" in v._fold:  # border-class weight sums of the folded inference path
                ops.call("b21_border_weight_sums", ops.ptr(v.w32), ops.ptr(v._fold["ws"]), v.cout, v.cin_true, v.taps,
                         ops.stream_ptr())
        for dst, src in self._aliases:
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src.detach().reshape(dst.shape))

    def _pc(self, name: str, conv: nn.Conv3d, cin_padded: Optional[int] = None):
        self._packed[name] = ops.PackedConv(conv.weight, conv.bias, cin_padded=cin_padded)

    def _vec(self, name: str, t: torch.Tensor):
        v = t.detach().reshape(-1).to(torch.float32).contiguous()
        self._packed[name] = v
        self._aliases.append((v, t))

    def _mat(self, name: str, t: torch.Tensor):
        v = t.detach().to(torch.float32).reshape(t.shape[0], -1).contiguous()
        self._packed[name] = v
        self._aliases.append((v, t))

    # ---- workspaces
    def _buf(self, ws, name, shape, dtype=torch.bfloat16):
        t = ws.get(name)
        if t is None or t.shape != tuple(shape) or t.dtype != dtype:
            t = torch.empty(tuple(shape), dtype=dtype, device=self._device())
            ws[name] = t
        return t

    def _device(self):
        return next(self.parameters()).device

    def _check_input(self, x):
        if not x.is_cuda:
            raise RuntimeError("brats21_b200 networks run on CUDA only (no CPU fallback): move the input to a B200")
        if x.dim() != 5 or x.shape[1] != self.inplanes:
            raise ValueError(f"expected [N, {self.inplanes}, D, H, W], got {tuple(x.shape)}")
        if any(s % 8 for s in x.shape[2:]):
            raise ValueError(f"spatial dims must be divisible by 8 (shape_to_divisible), got {tuple(x.shape[2:])}")
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters())) \
                and self.training:
            from .autograd import network_forward_train  # noqa: WPS433 (lazy: training path)
            return network_forward_train
        return None

    def pack_input(self, x: torch.Tensor) -> torch.Tensor:
        """NCDHW fp32 -> channels-last bf16 with the 4 modalities padded to 8 channels."""
        n, c, d, h, w = x.shape
        x = x.detach().to(torch.float32).contiguous()
        ws = self._ws.setdefault(("in", n, d, h, w), {})
        out = self._buf(ws, "x8", (n, d, h, w, 8))
        ops.pack_windows(x, out, [(0, 0, 0)] * n, vol_index=list(range(n)))
        return out

    def forward_infer(self, x8: torch.Tensor) -> torch.Tensor:
        """Logits of forward_packed(x8, want_deep=False).  The ~300 launches of one forward are captured once per
        (input buffer, shape, weights) into a CUDA graph and replayed: the sliding-window loop re-uses the same window
        buffer for every batch, so each further batch costs one graph launch instead of ~300 kernel launches."""
        self._ensure_packed()
        if not ops.use_graphs or ops.conv_profile is not None:
            return self.forward_packed(x8, want_deep=False)[0]
        key = (x8.data_ptr(), tuple(x8.shape), tuple(x8.stride()), ops.use_fold, ops.use_march, ops.use_slide,
               ops.use_point)
        entry = self._graphs.get(key)
        if entry is None:
            self.forward_packed(x8, want_deep=False)  # warm-up: workspaces, lazily packed/folded weights, attributes
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            l0 = ops._lib.launch_count
            with torch.cuda.graph(graph):
                out = self.forward_packed(x8, want_deep=False)[0]
            entry = (graph, out, ops._lib.launch_count - l0)
            self._graphs[key] = entry
        entry[0].replay()
        ops._lib.launch_count += entry[2]
        return entry[1]

    def forward(self, x: torch.Tensor):
        train_fn = self._check_input(x)
        if train_fn is not None:
            return train_fn(self, x)
        with torch.no_grad():
            want_deep = self.deep_supervision and not (self.skip_deep_heads_in_eval and not self.training)
            out, deeps = self.forward_packed(self.pack_input(x), want_deep)
        if self.deep_supervision:
            return out, deeps
        return out


# ================================================================================================ V1
class EquiUnet(_B21Net):
    """B200 EquiUnet (reference: networks/equiunet2020.py:408-500)."""
    name = "EquiUnet"

    def __init__(self, inplanes, num_classes, features, norm_layer=None, act="relu", deep_supervision=False,
                 dropout=0, refinement=False):
        super().__init__()
        if norm_layer == "bcn":
            raise NotImplementedError("norm_layer='bcn' (BCNorm, factory.py:128-176) is not on the accelerated path")
        if norm_layer not in _NORMS:
            raise ValueError("Norm type is not correct")  # get_norm_layer, factory.py:192
        if act not in _ACTS:
            raise NotImplementedError(f"act={act!r}: only relu / leakyrelu / elu are on the accelerated path")
        if refinement:
            raise NotImplementedError("refinement (RefUnet) is out of scope (SURVEY.md §0)")
        if dropout:
            raise NotImplementedError("dropout > 0 is not supported")
        f = list(features)
        self.inplanes, self.num_classes, self.features = inplanes, num_classes, f
        self.deep_supervision, self.act, self.refinement = deep_supervision, act, refinement
        self.norm = norm_layer
        kw = dict(norm=norm_layer)
        self.encoder1 = UBlock(inplanes, f[0], f[0], act, **kw)
        self.encoder2 = UBlock(f[0], f[1], f[1], act, **kw)
        self.encoder3 = UBlock(f[1], f[2], f[2], act, **kw)
        self.encoder4 = UBlock(f[2], f[3], f[3], act, **kw)
        self.bottom = UBlock(f[3], f[3], f[3], act, (2, 2), **kw)
        self.bottom_2 = ConvBnRelu(f[3] * 2, f[2], act, **kw)
        self.downsample = nn.MaxPool3d(2, 2)
        self.decoder3 = UBlock(f[2] * 2, f[2], f[1], act, **kw)
        self.decoder2 = UBlock(f[1] * 2, f[1], f[0], act, **kw)
        self.decoder1 = UBlock(f[0] * 2, f[0], f[0], act, **kw)
        self.upsample = nn.Upsample(scale_factor=2, mode="trilinear", align_corners=True)
        self.outconv = _conv1(f[0], num_classes)
        if deep_supervision:
            self.deep_bottom = _head(f[3], num_classes, 8)
            self.deep_bottom2 = _head(f[2], num_classes, 8)
            self.deep3 = _head(f[1], num_classes, 4)
            self.deep2 = _head(f[0], num_classes, 2)
        # init_weights(self, "kaiming") (factory.py:203-224): kaiming-normal fan_out on every Conv weight (biases keep
        # torch's default), N(1, 0.02) / 0 on BatchNorm weight / bias — drawn in module order, as net.apply() does
        for m in self.modules():
            if isinstance(m, nn.Conv3d):
                nn.init.kaiming_normal_(m.weight.data, a=0.0, mode="fan_out")
            elif isinstance(m, nn.BatchNorm3d):
                nn.init.normal_(m.weight.data, 1.0, 0.02)
                nn.init.constant_(m.bias.data, 0.0)
        self._init_runtime()

    _CBR = ["encoder1.ConvBnRelu1", "encoder1.ConvBnRelu2", "encoder2.ConvBnRelu1", "encoder2.ConvBnRelu2",
            "encoder3.ConvBnRelu1", "encoder3.ConvBnRelu2", "encoder4.ConvBnRelu1", "encoder4.ConvBnRelu2",
            "bottom.ConvBnRelu1", "bottom.ConvBnRelu2", "bottom_2", "decoder3.ConvBnRelu1", "decoder3.ConvBnRelu2",
            "decoder2.ConvBnRelu1", "decoder2.ConvBnRelu2", "decoder1.ConvBnRelu1", "decoder1.ConvBnRelu2"]

    def _pack(self):
        for name in self._CBR:
            m = self.get_submodule(name)
            self._pc(name, m.conv, cin_padded=8 if name == "encoder1.ConvBnRelu1" else None)
            if self.norm != "none":
                self._vec(name + ".g", m.bn.weight)
                self._vec(name + ".b", m.bn.bias)
        heads = ["outconv"] + (["deep_bottom.0", "deep_bottom2.0", "deep3.0", "deep2.0"] if self.deep_supervision else [])
        for name in heads:
            m = self.get_submodule(name)
            self._mat(name + ".w", m.weight)
            self._vec(name + ".bias", m.bias)

    def _recipe(self) -> bool:
        """GroupNorm(8) + ReLU — the README recipe — has the dedicated fused normalisation kernel and the training path."""
        return self.norm == "group" and self.act == "relu"

    def _grad_order(self):
        from .autograd import v1_grad_order
        return v1_grad_order(self)

    def _forward_train(self, x8, want_deep):
        if not self._recipe():
            raise NotImplementedError(f"EquiUnet(norm_layer={self.norm!r}, act={self.act!r}): the accelerated BACKWARD "
                                      "covers GroupNorm(8) + ReLU (the reference recipes, README.md:103-121); other "
                                      "factory combinations run forward / inference only")
        from .autograd import _v1_forward_train
        return _v1_forward_train(self, x8, want_deep)

    def _backward_train(self, tape, dout, ddeeps, gs):
        from .autograd import _backward_v1
        return _backward_v1(self, tape, dout, ddeeps, gs)

    def _cbr(self, name, x, out, stats, dil=1):
        p = self._packed
        if self._recipe():
            ops.conv3d(x, p[name], out=out, stats=stats, dil=dil)
            ops.norm_apply(out, stats, p[name + ".g"], p[name + ".b"], ops.GN_RELU)
            return out
        # the rest of the factory (networks/factory.py:179-200): statistics -> per-(n, c) affine -> activation
        if self.norm == "none":
            ops.conv3d(x, p[name], out=out, dil=dil)  # the conv carries the bias (equiunet2020.py:68)
            return ops.norm_act(out, ops.NORM_NONE, self.act)
        if self.norm == "group":
            ops.conv3d(x, p[name], out=out, stats=stats, dil=dil)
            return ops.norm_act(out, ops.NORM_GROUP, self.act, p[name + ".g"], p[name + ".b"], conv_stats=stats)
        ops.conv3d(x, p[name], out=out, dil=dil)
        if self.norm == "instance":
            return ops.norm_act(out, ops.NORM_INSTANCE, self.act, p[name + ".g"], p[name + ".b"])
        bn = self.get_submodule(name).bn
        if self.training:  # batch statistics; running statistics updated in place, as nn.BatchNorm3d.forward does
            bn.num_batches_tracked += 1
            return ops.norm_act(out, ops.NORM_BATCH_TRAIN, self.act, p[name + ".g"], p[name + ".b"],
                                running=(bn.running_mean, bn.running_var), momentum=bn.momentum)
        return ops.norm_act(out, ops.NORM_BATCH_EVAL, self.act, p[name + ".g"], p[name + ".b"],
                            running=(bn.running_mean, bn.running_var))

    def forward_packed(self, x8: torch.Tensor, want_deep: bool = True):
        """x8: [N, D, H, W, 8] bf16 channels-last (modalities in channels 0..3). Returns (logits, [deep heads])."""
        self._ensure_packed()
        n, d, h, w, _ = x8.shape
        f = self.features
        ws = self._ws.setdefault(("v1", n, d, h, w), {})
        B = lambda name, s, c: self._buf(ws, name, (n, d // s, h // s, w // s, c))  # noqa: E731
        stats = self._buf(ws, "stats", (ops._lib.STAT_SLOTS, n, 8, 2), torch.float64)
        cat1, cat2, cat3, cat4 = B("cat1", 1, 2 * f[0]), B("cat2", 2, 2 * f[1]), B("cat3", 4, 2 * f[2]), B("cat4", 8, 2 * f[3])
        t1, t2, t3, t4 = B("t1", 1, f[0]), B("t2", 2, f[1]), B("t3", 4, f[2]), B("t4", 8, f[3])
        p1, p2, p3 = B("p1", 2, f[0]), B("p2", 4, f[1]), B("p3", 8, f[2])

        self._cbr("encoder1.ConvBnRelu1", x8, t1, stats)
        down1 = self._cbr("encoder1.ConvBnRelu2", t1, cat1[..., :f[0]], stats)
        ops.scale_pool(down1, pooled=p1, mode=1)
        self._cbr("encoder2.ConvBnRelu1", p1, t2, stats)
        down2 = self._cbr("encoder2.ConvBnRelu2", t2, cat2[..., :f[1]], stats)
        ops.scale_pool(down2, pooled=p2, mode=1)
        self._cbr("encoder3.ConvBnRelu1", p2, t3, stats)
        down3 = self._cbr("encoder3.ConvBnRelu2", t3, cat3[..., :f[2]], stats)
        ops.scale_pool(down3, pooled=p3, mode=1)
        self._cbr("encoder4.ConvBnRelu1", p3, t4, stats)
        down4 = self._cbr("encoder4.ConvBnRelu2", t4, cat4[..., :f[3]], stats)
        self._cbr("bottom.ConvBnRelu1", down4, t4, stats, dil=2)
        bottom = self._cbr("bottom.ConvBnRelu2", t4, cat4[..., f[3]:], stats, dil=2)
        b2 = self._cbr("bottom_2", cat4, B("b2", 8, f[2]), stats)
        ops.upsample2x(b2, cat3[..., f[2]:])
        self._cbr("decoder3.ConvBnRelu1", cat3, t3, stats)
        u3 = self._cbr("decoder3.ConvBnRelu2", t3, B("u3", 4, f[1]), stats)
        ops.upsample2x(u3, cat2[..., f[1]:])
        self._cbr("decoder2.ConvBnRelu1", cat2, t2, stats)
        u2 = self._cbr("decoder2.ConvBnRelu2", t2, B("u2", 2, f[0]), stats)
        ops.upsample2x(u2, cat1[..., f[0]:])
        self._cbr("decoder1.ConvBnRelu1", cat1, t1, stats)
        u1 = self._cbr("decoder1.ConvBnRelu2", t1, B("u1", 1, f[0]), stats)
        p = self._packed
        out = ops.head_conv(u1, p["outconv.w"], p["outconv.bias"])
        deeps: List[torch.Tensor] = []
        if want_deep and self.deep_supervision:
            for name, src, s in (("deep_bottom.0", bottom, 8), ("deep_bottom2.0", b2, 8), ("deep3.0", u3, 4),
                                 ("deep2.0", u2, 2)):
                deeps.append(ops.upsample_f32(ops.head_conv(src, p[name + ".w"], p[name + ".bias"]), s))
        return out, deeps


# ================================================================================================ V2
class EquiUnetASSPEvo(_B21Net):
    """B200 EquiUnet-ASPP-Evo (reference: networks/equiunet2021.py:225-333)."""
    name = "EquiUnetASSPEvo"

    def __init__(self, inplanes, num_classes, features, norm_layer=None, act="relu", deep_supervision=False,
                 dropout=0, refinement=False):
        super().__init__()
        warnings.warn("norm layer and activation specified will not be used ! only EVO !!")  # as the reference
        if refinement:
            raise NotImplementedError("refinement is broken in the reference (equiunet2021.py:237,282) — out of scope")
        if dropout:
            raise NotImplementedError("dropout > 0 is not supported")
        f = list(features)
        if f[0] % 16:
            raise ValueError("EvoNorm groups=8 on width/2 channels needs width to be a multiple of 16")
        self.inplanes, self.num_classes, self.features = inplanes, num_classes, f
        self.deep_supervision, self.act, self.refinement = deep_supervision, act.upper(), refinement
        self.encoder1 = ConvEvoBlockCorrected(inplanes, f[0], dropout)
        self.encoder2 = ConvEvoBlockCorrected(2 * f[0], f[1], dropout)
        self.encoder3 = ConvEvoBlockCorrected(2 * f[1], f[2], dropout)
        self.encoder4 = ConvEvoBlockCorrected(2 * f[2], f[3], dropout)
        self.bridge1 = ConvEvo(f[0], f[0] // 2, dropout)
        self.bridge2 = ConvEvo(f[1], f[1] // 2, dropout)
        self.bridge3 = ConvEvo(f[2], f[2] // 2, dropout)
        self.aspp = SimpleASPPEVO(f[3], f[3] // 4)
        self.downsample = nn.Identity()  # MONAI MaxAvgPool has no parameters; fused into scale_pool
        self.upconv3 = ConvEvo(f[3], f[3] // 4, dropout)
        self.decoder3 = ConvEvoBlockCorrected(f[2], f[2], dropout)
        self.upconv2 = ConvEvo(f[2], f[2] // 4, dropout)
        self.decoder2 = ConvEvoBlockCorrected(f[1], f[1], dropout)
        self.upconv1 = ConvEvo(f[1], f[1] // 4, dropout)
        self.decoder1 = ConvEvoBlockCorrected(f[0], f[0], dropout)
        self.upsample = nn.Upsample(scale_factor=2, mode="trilinear", align_corners=True)
        self.out_conv = _conv1(f[0], num_classes)
        if deep_supervision:
            self.deep3 = _head(f[2], num_classes, 4)
            self.deep2 = _head(f[1], num_classes, 2)
        self._init_runtime()

    _BLOCKS = ["encoder1", "encoder2", "encoder3", "encoder4", "decoder3", "decoder2", "decoder1"]
    _CONVEVO = ["bridge1", "bridge2", "bridge3", "aspp.conv_k1", "upconv3", "upconv2", "upconv1"]

    def _pack(self):
        for b in self._BLOCKS:
            seq = self.get_submodule(b).conv_conv_se
            self._pc(b + ".c0", seq[0], cin_padded=8 if b == "encoder1" else None)
            self._pc(b + ".c1", seq[3])
            for tag, evo in (("e0", seq[1]), ("e1", seq[4])):
                self._vec(f"{b}.{tag}.g", evo.gamma)
                self._vec(f"{b}.{tag}.b", evo.beta)
            fc = seq[6].fc
            self._mat(b + ".se.w1", fc[0].weight)
            self._vec(b + ".se.b1", fc[0].bias)
            self._mat(b + ".se.w2", fc[2].weight)
            self._vec(b + ".se.b2", fc[2].bias)
        for name in self._CONVEVO:
            m = self.get_submodule(name)
            self._pc(name, m.conv)
            self._vec(name + ".g", m.evo.gamma)
            self._vec(name + ".b", m.evo.beta)
        for i, conv in enumerate(self.aspp.convs):
            self._pc(f"aspp.convs.{i}", conv)
        heads = ["out_conv"] + (["deep3.0", "deep2.0"] if self.deep_supervision else [])
        for name in heads:
            m = self.get_submodule(name)
            self._mat(name + ".w", m.weight)
            self._vec(name + ".bias", m.bias)

    def _grad_order(self):
        from .autograd import v2_grad_order
        return v2_grad_order(self)

    def _forward_train(self, x8, want_deep):
        from .autograd import _v2_forward_train
        return _v2_forward_train(self, x8, want_deep)

    def _backward_train(self, tape, dout, ddeeps, gs):
        from .autograd import _backward_v2
        return _backward_v2(self, tape, dout, ddeeps, gs)

    def _block(self, name, x, tmp, out, stats, csum):
        """conv-evo-conv-evo + SE gate; returns (pre-scale activations in `out`, scale [N, C])."""
        p = self._packed
        ops.conv3d(x, p[name + ".c0"], out=tmp, stats=stats)
        ops.norm_apply(tmp, stats, p[name + ".e0.g"], p[name + ".e0.b"], ops.EVO_S0)
        ops.conv3d(tmp, p[name + ".c1"], out=out, stats=stats)
        csum.zero_()
        ops.norm_apply(out, stats, p[name + ".e1.g"], p[name + ".e1.b"], ops.EVO_S0, chan_sum=csum)
        n, d, h, w, _ = out.shape
        scale = ops.se_gate(csum, p[name + ".se.w1"], p[name + ".se.b1"], p[name + ".se.w2"], p[name + ".se.b2"],
                            d * h * w)
        return out, scale

    def _convevo(self, name, x, out, stats):
        p = self._packed
        ops.conv3d(x, p[name], out=out, stats=stats)
        ops.norm_apply(out, stats, p[name + ".g"], p[name + ".b"], ops.EVO_S0)
        return out

    # ---- folded-EvoNorm inference path (csrc/fold.cu): levels 1-2 never run a normalisation pass
    _FOLD_CONVS = ["encoder1.c0", "encoder1.c1", "encoder2.c0", "encoder2.c1", "decoder2.c0", "decoder2.c1",
                   "decoder1.c0", "decoder1.c1", "bridge1", "bridge2", "upconv1", "upconv2"]

    _FOLD3_CONVS = ["encoder3.c0", "encoder3.c1", "decoder3.c0", "decoder3.c1", "bridge3"]

    def _fold_ok(self, d, h, w):
        return ops.use_fold and min(h, w) >= 16 and d >= 4 and all(ops.fold_supported(self._packed[k])
                                                                    for k in self._FOLD_CONVS)

    def _ab(self, ws, name, n, c):
        t = self._buf(ws, name, (2, n, c), torch.float32)
        return t[0], t[1]

    def _fblock(self, name, x, xab, tmp, out, stats, csum, ab_tmp, ab_out):
        """ConvEvoBlockCorrected on stored-swish tensors: conv -> S0 (+stats) ; conv(fold a0,b0) -> S1 (+stats, channel
        sums) ; (A, B) of the block output = EvoNorm affine x SE gate."""
        p = self._packed
        n, d, h, w, _ = out.shape
        nvox = d * h * w
        ops.conv3d_fold(x, p[name + ".c0"], tmp, stats, ab=xab, act=True)
        ops.evo_se_affine(stats, p[name + ".e0.g"], p[name + ".e0.b"], ab_tmp[0], ab_tmp[1], nvox)
        csum.zero_()
        ops.conv3d_fold(tmp, p[name + ".c1"], out, stats, ab=ab_tmp, act=True, chan_sum=csum)
        ops.evo_se_affine(stats, p[name + ".e1.g"], p[name + ".e1.b"], ab_out[0], ab_out[1], nvox, chan_sum=csum,
                          se=(p[name + ".se.w1"], p[name + ".se.b1"], p[name + ".se.w2"], p[name + ".se.b2"]))

    def _fconvevo(self, name, x, xab, out, stats, ab_out):
        """ConvEvo (1x1) on a stored-swish input; output stored as swish, its EvoNorm affine written to ab_out."""
        p = self._packed
        n, d, h, w, _ = out.shape
        ops.conv3d_fold(x, p[name], out, stats, ab=xab, act=True)
        ops.evo_se_affine(stats, p[name + ".g"], p[name + ".b"], ab_out[0], ab_out[1], d * h * w)

    def forward_packed_folded(self, x8: torch.Tensor, want_deep: bool = True):
        n, d, h, w, _ = x8.shape
        f = self.features
        ws = self._ws.setdefault(("v2f", n, d, h, w), {})
        B = lambda name, s, c: self._buf(ws, name, (n, d // s, h // s, w // s, c))  # noqa: E731
        stats = self._buf(ws, "stats", (ops._lib.STAT_SLOTS, n, 8, 2), torch.float64)
        cs = [self._buf(ws, f"csum{i}", (n, f[i]), torch.float32) for i in range(4)]
        t1, t2, t3, t4 = B("t1", 1, f[0]), B("t2", 2, f[1]), B("t3", 4, f[2]), B("t4", 8, f[3])
        d1, d2, d3, d4 = B("d1", 1, f[0]), B("d2", 2, f[1]), B("d3", 4, f[2]), B("d4", 8, f[3])
        p1, p2, p3 = B("p1", 2, 2 * f[0]), B("p2", 4, 2 * f[1]), B("p3", 8, 2 * f[2])
        cat1, cat2, cat3 = None, B("cat2", 2, f[1]), B("cat3", 4, f[2])
        ab_t1, ab_t2 = self._ab(ws, "ab_t1", n, f[0]), self._ab(ws, "ab_t2", n, f[1])
        ab_d1, ab_d2 = self._ab(ws, "ab_d1", n, f[0]), self._ab(ws, "ab_d2", n, f[1])
        ab_c1, ab_c2 = self._ab(ws, "ab_c1", n, f[0]), self._ab(ws, "ab_c2", n, f[1])
        ab_u1, ab_u2 = self._ab(ws, "ab_u1", n, f[0]), self._ab(ws, "ab_u2", n, f[1])
        h0, h1 = f[0] // 2, f[1] // 2

        # levels 1-2 of the encoder: stored-swish tensors, pooled outputs are actual values
        self._fblock("encoder1", x8, None, t1, d1, stats, cs[0], ab_t1, ab_d1)
        ops.affine_pool(d1, ab_d1[0], ab_d1[1], p1, mode=2)
        self._fblock("encoder2", p1, None, t2, d2, stats, cs[1], ab_t2, ab_d2)
        ops.affine_pool(d2, ab_d2[0], ab_d2[1], p2, mode=2)
        # level 3 is folded as well when its planes are large enough for the sliding-window kernel (it removes five
        # of the nine normalisation passes, two SE-gate launches and two scale passes per batch); level 4 (tap kernel,
        # 16^3) keeps the explicit normalisation
        h2 = f[2] // 2
        fold3 = ops.fold_level3 and d // 4 >= 2 and h // 4 >= 8 and w // 4 >= 8 and \
            all(ops.fold_supported(self._packed[k]) for k in self._FOLD3_CONVS)
        if fold3:
            ab_t3, ab_d3 = self._ab(ws, "ab_t3", n, f[2]), self._ab(ws, "ab_d3", n, f[2])
            ab_u3 = self._ab(ws, "ab_u3", n, f[2])
            fresh = "ab_c3" not in ws
            ab_c3 = self._ab(ws, "ab_c3", n, f[2])
            if fresh:  # the up-sampled half of the level-3 concat holds actual values: identity affine, set once
                ab_c3[0][:, h2:] = 1.0
                ab_c3[1][:, h2:] = 0.0
            self._fblock("encoder3", p2, None, t3, d3, stats, cs[2], ab_t3, ab_d3)
            ops.affine_pool(d3, ab_d3[0], ab_d3[1], p3, mode=2)
        else:
            _, s = self._block("encoder3", p2, t3, d3, stats, cs[2])
            ops.scale_pool(d3, s, full=d3, pooled=p3, mode=2)
        _, s = self._block("encoder4", p3, t4, d4, stats, cs[3])
        ops.scale_pool(d4, s, full=d4, mode=0)
        pk = self._packed
        acat = B("asppcat", 8, f[3])
        q = f[3] // 4
        for i, dil in enumerate(self.aspp.dilations):
            ops.conv3d(d4, pk[f"aspp.convs.{i}"], out=acat[..., i * q:(i + 1) * q], dil=dil)
        assp = self._convevo("aspp.conv_k1", acat, t4, stats)

        # level-1 concat [bridge1 | up(upconv1)]: two DENSE 24-channel tensors read side by side by decoder1's first
        # conv (a 48-byte half of a 96-byte record is a partial-sector write: 1.3x the DRAM traffic on both producers)
        split1 = ops.split_concat and self._packed["decoder1.c0"].w_march is not None
        if split1:
            cat1a, cat1b = B("cat1a", 1, h0), B("cat1b", 1, h0)
        else:
            cat1 = B("cat1", 1, f[0])
            cat1a, cat1b = cat1[..., :h0], cat1[..., h0:]
        self._fconvevo("bridge1", d1, ab_d1, cat1a, stats, (ab_c1[0][:, :h0], ab_c1[1][:, :h0]))
        self._fconvevo("bridge2", d2, ab_d2, cat2[..., :h1], stats, (ab_c2[0][:, :h1], ab_c2[1][:, :h1]))
        if fold3:
            self._fconvevo("bridge3", d3, ab_d3, cat3[..., :h2], stats, (ab_c3[0][:, :h2], ab_c3[1][:, :h2]))
        else:
            self._convevo("bridge3", d3, cat3[..., :h2], stats)

        u = self._convevo("upconv3", assp, B("uc3", 8, f[3] // 4), stats)
        ops.upsample2x(u, cat3[..., h2:])
        up3 = B("up3", 4, f[2])
        if fold3:
            self._fblock("decoder3", cat3, ab_c3, t3, up3, stats, cs[2], ab_t3, ab_u3)
        else:
            _, s = self._block("decoder3", cat3, t3, up3, stats, cs[2])
            ops.scale_pool(up3, s, full=up3, mode=0)
        uc2 = B("uc2", 4, f[2] // 4)
        self._fconvevo("upconv2", up3, ab_u3 if fold3 else None, uc2, stats, (ab_c2[0][:, h1:], ab_c2[1][:, h1:]))
        ops.upsample2x(uc2, cat2[..., h1:])  # interpolation weights sum to 1: the affine passes through unchanged
        up2 = B("up2", 2, f[1])
        self._fblock("decoder2", cat2, ab_c2, t2, up2, stats, cs[1], ab_t2, ab_u2)
        uc1 = B("uc1", 2, f[1] // 4)
        self._fconvevo("upconv1", up2, ab_u2, uc1, stats, (ab_c1[0][:, h0:], ab_c1[1][:, h0:]))
        ops.upsample2x(uc1, cat1b)
        up1 = B("up1", 1, f[0])
        self._fblock("decoder1", (cat1a, cat1b) if split1 else cat1, ab_c1, t1, up1, stats, cs[0], ab_t1, ab_u1)
        out = ops.head_conv(up1, pk["out_conv.w"], pk["out_conv.bias"], scale=ab_u1[0], offset=ab_u1[1])
        deeps: List[torch.Tensor] = []
        if want_deep and self.deep_supervision:
            if fold3:
                deeps.append(ops.upsample_f32(ops.head_conv(up3, pk["deep3.0.w"], pk["deep3.0.bias"], scale=ab_u3[0],
                                                            offset=ab_u3[1]), 4))
            else:
                deeps.append(ops.upsample_f32(ops.head_conv(up3, pk["deep3.0.w"], pk["deep3.0.bias"]), 4))
            deeps.append(ops.upsample_f32(ops.head_conv(up2, pk["deep2.0.w"], pk["deep2.0.bias"], scale=ab_u2[0],
                                                        offset=ab_u2[1]), 2))
        return out, deeps

    def forward_packed(self, x8: torch.Tensor, want_deep: bool = True):
        self._ensure_packed()
        n, d, h, w, _ = x8.shape
        if self._fold_ok(d, h, w):
            return self.forward_packed_folded(x8, want_deep)
        f = self.features
        ws = self._ws.setdefault(("v2", n, d, h, w), {})
        B = lambda name, s, c: self._buf(ws, name, (n, d // s, h // s, w // s, c))  # noqa: E731
        stats = self._buf(ws, "stats", (ops._lib.STAT_SLOTS, n, 8, 2), torch.float64)
        cs = [self._buf(ws, f"csum{i}", (n, f[i]), torch.float32) for i in range(4)]
        t1, t2, t3, t4 = B("t1", 1, f[0]), B("t2", 2, f[1]), B("t3", 4, f[2]), B("t4", 8, f[3])
        d1, d2, d3, d4 = B("d1", 1, f[0]), B("d2", 2, f[1]), B("d3", 4, f[2]), B("d4", 8, f[3])
        p1, p2, p3 = B("p1", 2, 2 * f[0]), B("p2", 4, 2 * f[1]), B("p3", 8, 2 * f[2])
        cat1, cat2, cat3 = B("cat1", 1, f[0]), B("cat2", 2, f[1]), B("cat3", 4, f[2])

        _, s = self._block("encoder1", x8, t1, d1, stats, cs[0])
        ops.scale_pool(d1, s, full=d1, pooled=p1, mode=2)
        _, s = self._block("encoder2", p1, t2, d2, stats, cs[1])
        ops.scale_pool(d2, s, full=d2, pooled=p2, mode=2)
        _, s = self._block("encoder3", p2, t3, d3, stats, cs[2])
        ops.scale_pool(d3, s, full=d3, pooled=p3, mode=2)
        _, s = self._block("encoder4", p3, t4, d4, stats, cs[3])
        ops.scale_pool(d4, s, full=d4, mode=0)

        pk = self._packed
        acat = B("asppcat", 8, f[3])
        q = f[3] // 4
        for i, dil in enumerate(self.aspp.dilations):
            ops.conv3d(d4, pk[f"aspp.convs.{i}"], out=acat[..., i * q:(i + 1) * q], dil=dil)
        assp = self._convevo("aspp.conv_k1", acat, t4, stats)

        self._convevo("bridge1", d1, cat1[..., :f[0] // 2], stats)
        self._convevo("bridge2", d2, cat2[..., :f[1] // 2], stats)
        self._convevo("bridge3", d3, cat3[..., :f[2] // 2], stats)

        u = self._convevo("upconv3", assp, B("uc3", 8, f[3] // 4), stats)
        ops.upsample2x(u, cat3[..., f[2] // 2:])
        up3, s = self._block("decoder3", cat3, t3, B("up3", 4, f[2]), stats, cs[2])
        ops.scale_pool(up3, s, full=up3, mode=0)
        u = self._convevo("upconv2", up3, B("uc2", 4, f[2] // 4), stats)
        ops.upsample2x(u, cat2[..., f[1] // 2:])
        up2, s = self._block("decoder2", cat2, t2, B("up2", 2, f[1]), stats, cs[1])
        ops.scale_pool(up2, s, full=up2, mode=0)
        u = self._convevo("upconv1", up2, B("uc1", 2, f[1] // 4), stats)
        ops.upsample2x(u, cat1[..., f[0] // 2:])
        up1, s1 = self._block("decoder1", cat1, t1, B("up1", 1, f[0]), stats, cs[0])
        # the SE scale of decoder1 is folded into the 1x1 head: W (x * s) == (W * s) x
        out = ops.head_conv(up1, pk["out_conv.w"], pk["out_conv.bias"], scale=s1)
        deeps: List[torch.Tensor] = []
        if want_deep and self.deep_supervision:
            deeps.append(ops.upsample_f32(ops.head_conv(up3, pk["deep3.0.w"], pk["deep3.0.bias"]), 4))
            deeps.append(ops.upsample_f32(ops.head_conv(up2, pk["deep2.0.w"], pk["deep2.0.bias"]), 2))
        return out, deeps
