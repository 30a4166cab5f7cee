"""Python-side op wrappers over the C-ABI (libb21.so).  Activations are channels-last bf16 tensors
[N, D, H, W, C]; a tensor may be a channel slice (`t[..., a:b]`) of a wider buffer — only the last-dim stride
must be 1 and the voxel stride (`t.stride(3)`) is passed as the leading dimension.
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import call, ptr, stream_ptr


def _ld(t: torch.Tensor) -> int:
    """Channel stride (elements between consecutive voxels) of a dense-voxel channels-last tensor."""
    n, d, h, w, c = t.shape
    ld = t.stride(3)
    assert t.stride(4) == 1 and t.stride(2) == w * ld and t.stride(1) == h * w * ld and \
        (n == 1 or t.stride(0) == d * h * w * ld), f"not a channels-last voxel-dense view: {t.shape} {t.stride()}"
    return ld


class PackedConv:
    """bf16 weight repacked for the implicit-GEMM kernels: [k^3][cout_padded][cin_padded] (+ fp32 bias)."""

    def __init__(self, weight: torch.Tensor, bias, cin_padded: int | None = None, transpose_flip: bool = False):
        assert weight.is_cuda and weight.dim() == 5
        cout, cin, k = weight.shape[0], weight.shape[1], weight.shape[2]
        rows, inner = (cin, cout) if transpose_flip else (cout, cin)
        if cin_padded is None:
            cin_padded = (inner + 7) // 8 * 8
        self.k, self.taps = k, k ** 3
        self.cout, self.cin = rows, cin_padded
        self.cout_padded = _lib.load().b21_conv_cout_padded(rows)
        self.w = torch.empty((self.taps, self.cout_padded, cin_padded), dtype=torch.bfloat16, device=weight.device)
        w32 = weight.detach().to(torch.float32).contiguous()
        call("b21_pack_conv_weight", ptr(w32), ptr(self.w), cout, cin, cin_padded, k, int(transpose_flip),
             stream_ptr())
        self.bias = None if bias is None else bias.detach().to(torch.float32).contiguous()


def new_stats(n: int, device) -> torch.Tensor:
    return torch.empty((_lib.STAT_SLOTS, n, 8, 2), dtype=torch.float64, device=device)


def conv3d(x: torch.Tensor, pw: PackedConv, out: torch.Tensor | None = None, stats: torch.Tensor | None = None,
           dil: int = 1) -> torch.Tensor:
    """Same-padded stride-1 conv3d (k = 1 or 3) on channels-last bf16; optional group statistics side output."""
    n, d, h, w, cin = x.shape
    assert x.dtype == torch.bfloat16 and cin == pw.cin, (x.dtype, cin, pw.cin)
    if out is None:
        out = torch.empty((n, d, h, w, pw.cout), dtype=torch.bfloat16, device=x.device)
    assert out.shape == (n, d, h, w, pw.cout) and out.dtype == torch.bfloat16
    call("b21_conv3d_fwd", ptr(x), _ld(x), ptr(pw.w), ptr(pw.bias), ptr(out), _ld(out), ptr(stats),
         n, d, h, w, cin, pw.cout, pw.taps, dil, stream_ptr())
    return out
