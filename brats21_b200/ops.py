"""Python-side op wrappers over the C-ABI (libb21.so).  Activations are channels-last bf16 tensors
[N, D, H, W, C]; a tensor may be a channel slice (`t[..., a:b]`) of a wider buffer — only the last-dim stride
must be 1 and the voxel stride (`t.stride(3)`) is passed as the leading dimension.
"""
from __future__ import annotations

import os

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


# When set to a list, every HBM-bound launch appends (start_event, end_event, algorithmic_bytes, kind): bench.py's
# `hbm` roofline block (SURVEY.md §8d byte counts: bf16 activations read once / written once, fp32 accumulators RMW).
hbm_profile = None


class _hbm:
    __slots__ = ("kind", "nbytes", "e0")

    def __init__(self, kind: str, nbytes: float):
        self.kind, self.nbytes, self.e0 = kind, nbytes, None

    def __enter__(self):
        if hbm_profile is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if self.e0 is not None and hbm_profile is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            hbm_profile.append((self.e0, e1, float(self.nbytes), self.kind))
        return False


class PackedConv:
    """bf16 weight repacked for the implicit-GEMM kernels: [k^3][cout_padded][cin_padded] (+ fp32 bias).

    ``w32`` / ``bias`` alias the live fp32 parameter storage whenever the parameter is already fp32 and contiguous, so
    after an in-place parameter update (optimizer step) ``refresh()`` re-runs the packing kernels INTO THE SAME
    BUFFERS: pointers captured by CUDA graphs, tensor maps and the per-sample fold buffers stay valid."""

    def __init__(self, weight: torch.Tensor, bias, cin_padded: int | None = None, transpose_flip: bool = False):
        assert weight.is_cuda and weight.dim() == 5
        cout, cin, k = weight.shape[0], weight.shape[1], weight.shape[2]
        rows, inner = (cin, cout) if transpose_flip else (cout, cin)
        if cin_padded is None:
            cin_padded = (inner + 7) // 8 * 8
        self.k, self.taps = k, k ** 3
        self.cout, self.cin, self.cin_true = rows, cin_padded, inner
        self.cout_padded = _lib.load().b21_conv_cout_padded(rows)
        self.w = torch.empty((self.taps, self.cout_padded, cin_padded), dtype=torch.bfloat16, device=weight.device)
        self._src_w, self._src_b, self._tf = weight, bias, bool(transpose_flip)
        self._wshape = (cout, cin)
        self.w32 = weight.detach().to(torch.float32).contiguous()
        self.cin_padded, self._fold = cin_padded, {}
        self.bias = None if bias is None else bias.detach().to(torch.float32).contiguous()
        lib = _lib.load()
        self.point_ok = k == 1 and bool(lib.b21_conv_point_supported(cin_padded, rows))
        # plane-marching packing (k = 3, weights resident in shared memory) when the shape allows it
        self.w_march = None
        if k == 3 and lib.b21_conv_march_supported(cin_padded, rows):
            nbytes = lib.b21_conv_march_weight_bytes(cin_padded, rows)
            self.w_march = torch.empty((nbytes // 2,), dtype=torch.bfloat16, device=weight.device)
        # sliding-window packing (k = 3, 16 <= cin <= 96, weights streamed per tap)
        self.w_slide = None
        if k == 3 and self.w_march is None and lib.b21_conv_slide_supported(cin_padded, rows):
            nbytes = lib.b21_conv_slide_weight_bytes(cin_padded, rows)
            self.w_slide = torch.empty((nbytes // 2,), dtype=torch.bfloat16, device=weight.device)
        # first-conv packing (k = 3, at most 4 real input channels): im2col plane blocks, see csrc/conv_input.cu
        self.w_input = None
        if k == 3 and not transpose_flip and lib.b21_conv_input_supported(inner, rows):
            nbytes = lib.b21_conv_input_weight_bytes(rows)
            self.w_input = torch.empty((nbytes // 2,), dtype=torch.bfloat16, device=weight.device)
        self._launch_packs()

    def _launch_packs(self):
        cout, cin = self._wshape
        tf = int(self._tf)
        call("b21_pack_conv_weight", ptr(self.w32), ptr(self.w), cout, cin, self.cin_padded, self.k, tf, stream_ptr())
        if self.w_march is not None:
            call("b21_pack_conv_weight_march", ptr(self.w32), ptr(self.w_march), cout, cin, tf, stream_ptr())
        if self.w_slide is not None:
            call("b21_pack_conv_weight_slide", ptr(self.w32), ptr(self.w_slide), cout, cin, tf, stream_ptr())
        if self.w_input is not None:
            call("b21_pack_conv_weight_input", ptr(self.w32), ptr(self.w_input), cout, cin, stream_ptr())
        if "ws" in self._fold:
            call("b21_border_weight_sums", ptr(self.w32), ptr(self._fold["ws"]), self.cout, self.cin_true, self.taps,
                 stream_ptr())

    def sync_sources(self):
        """Bring the fp32 views up to date when they are converted copies (parameters that are not fp32 / contiguous)."""
        if self.w32.data_ptr() != self._src_w.data_ptr():
            self.w32.copy_(self._src_w.detach())
        if self.bias is not None and self.bias.data_ptr() != self._src_b.data_ptr():
            self.bias.copy_(self._src_b.detach())

    def refresh(self):
        """Re-pack after the source parameters changed in place (same storage): buffers are re-used."""
        self.sync_sources()
        self._launch_packs()

    def pack_jobs(self):
        """Job records of b21_pack_batch that reproduce _launch_packs (without the fold's border sums)."""
        import ctypes as C
        cout, cin = self._wshape
        tf = int(self._tf)
        jobs = []

        def add(fn, *args):
            job = _lib.PackJob()
            call(fn, *args, C.addressof(job))
            jobs.append(job)

        add("b21_pack_job_tap", ptr(self.w32), ptr(self.w), cout, cin, self.cin_padded, self.k, tf)
        if self.w_march is not None:
            add("b21_pack_job_march", ptr(self.w32), ptr(self.w_march), cout, cin, tf)
        if self.w_slide is not None:
            add("b21_pack_job_slide", ptr(self.w32), ptr(self.w_slide), cout, cin, tf)
        if self.w_input is not None:
            add("b21_pack_job_input", ptr(self.w32), ptr(self.w_input), cout, cin)
        return jobs


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
    prof = conv_profile
    if prof is not None:
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
    kind = "tap"
    if use_point and pw.taps == 1 and pw.point_ok:
        kind = "point"
        call("b21_conv1x1_fwd", ptr(x), _ld(x), ptr(pw.w), ptr(pw.bias), ptr(out), _ld(out), ptr(stats),
             n, d * h * w, cin, pw.cout, stream_ptr())
    elif use_input and dil == 1 and pw.w_input is not None and _ld(x) == 8:
        kind = "input"
        call("b21_conv3d_input_fwd", ptr(x), _ld(x), ptr(pw.w_input), ptr(pw.bias), ptr(out), _ld(out), ptr(stats), 0,
             n, d, h, w, pw.cout, stream_ptr())
    elif use_march and dil == 1 and pw.w_march is not None and h >= 8 and w >= 8:
        kind = "march"
        call("b21_conv3d_march_fwd", ptr(x), _ld(x), ptr(pw.w_march), ptr(pw.bias), ptr(out), _ld(out), ptr(stats),
             n, d, h, w, cin, pw.cout, stream_ptr())
    elif use_slide and dil == 1 and pw.w_slide is not None and h >= 8 and w >= 8 and (cin <= 96 or h * w >= 1024):
        # (channel-chunked mode, cin >= 128: the 16^3 level has too few tiles to balance the persistent grid — measured
        # 754 vs 891 TFLOP/s for 384 -> 384 — and stays on the tap kernel)
        kind = "slide"
        call("b21_conv3d_slide_fwd", ptr(x), _ld(x), ptr(pw.w_slide), ptr(pw.bias), ptr(out), _ld(out), ptr(stats),
             n, d, h, w, cin, pw.cout, stream_ptr())
    else:
        call("b21_conv3d_fwd", ptr(x), _ld(x), ptr(pw.w), ptr(pw.bias), ptr(out), _ld(out), ptr(stats),
             n, d, h, w, cin, pw.cout, pw.taps, dil, stream_ptr())
    if prof is not None:
        e1.record()
        prof.append((e0, e1, 2.0 * n * d * h * w * pw.cin_true * pw.cout * pw.taps, (kind, cin, pw.cout, pw.taps, d)))
        if kind == "point" and hbm_profile is not None:
            hbm_profile.append((e0, e1, 2.0 * n * d * h * w * (cin + pw.cout), "conv1x1"))
    return out


# ---------------------------------------------------------------------------------------------- folded EvoNorm
def evo_se_affine(stats, gamma, beta, a_out, b_out, nvox, chan_sum=None, se=None, eps=1e-5):
    """(A, B)[n][c] of EvoNorm-S0 [+ ResidualSE gate] from conv-epilogue statistics (see csrc/fold.cu).
    a_out / b_out: fp32 [n, c] views (row stride = ld) that receive A and B.  se = (w1, b1, w2, b2)."""
    n, c = a_out.shape
    assert a_out.stride(1) == 1 and b_out.stride(1) == 1 and a_out.stride(0) == b_out.stride(0)
    w1 = b1 = w2 = b2 = None
    hidden = 0
    if chan_sum is not None:
        w1, b1, w2, b2 = se
        hidden = w1.shape[0]
    call("b21_evo_se_affine", ptr(stats), ptr(gamma), ptr(beta), ptr(chan_sum), ptr(w1), ptr(b1), ptr(w2), ptr(b2),
         ptr(a_out), ptr(b_out), a_out.stride(0), n, c, hidden, nvox, eps, stream_ptr())


def _fold_prepare(pw: "PackedConv", a_in, b_in):
    """Per-sample packed weights W * A[n][ci] and the bias table bias + (border-class tap sums of W) . B[n]."""
    n, cin = a_in.shape
    assert cin == pw.cin_true == pw.cin and a_in.stride(1) == 1 and b_in.stride(0) == a_in.stride(0)
    dev = pw.w32.device
    f = pw._fold
    ncls = 27 if pw.taps == 27 else 1
    if "ws" not in f:
        f["ws"] = torch.empty((ncls, cin, pw.cout), dtype=torch.float32, device=dev)
        call("b21_border_weight_sums", ptr(pw.w32), ptr(f["ws"]), pw.cout, cin, pw.taps, stream_ptr())
    key = ("buf", n)
    if key not in f:
        if pw.taps == 1:
            nbytes = pw.cout_padded * pw.cin * 2
        elif pw.w_march is not None:
            nbytes = pw.w_march.numel() * 2
        else:
            nbytes = pw.w_slide.numel() * 2
        # zero-filled once: the per-sample packing rewrites the valid elements only (padding rows / channels stay 0)
        f[key] = (torch.zeros((n, nbytes // 2), dtype=torch.bfloat16, device=dev),
                  torch.empty((n, ncls, pw.cout), dtype=torch.float32, device=dev))
    packed, table = f[key]
    ld = a_in.stride(0)
    tail = (ptr(f["ws"]), ptr(pw.bias), ptr(b_in), ptr(table), stream_ptr())  # bias table in the same launch
    if pw.taps == 1:
        call("b21_pack_conv_weight_fold", ptr(pw.w32), ptr(packed), pw.cout, cin, pw.cin, 1, ptr(a_in), ld, n, *tail)
    elif pw.w_march is not None:
        call("b21_pack_conv_weight_march_fold", ptr(pw.w32), ptr(packed), pw.cout, cin, ptr(a_in), ld, n, *tail)
    else:
        call("b21_pack_conv_weight_slide_fold", ptr(pw.w32), ptr(packed), pw.cout, cin, ptr(a_in), ld, n, *tail)
    return packed, packed.stride(0) * 2, table


def fold_supported(pw: "PackedConv") -> bool:
    """True when conv3d_fold can run this conv (persistent 1x1, march or slide kernel)."""
    return (pw.taps == 1 and pw.point_ok) or (pw.taps == 27 and (pw.w_march is not None or pw.w_slide is not None))


def conv3d_fold(x, pw: "PackedConv", out, stats, ab=None, act=True, chan_sum=None):
    """conv3d whose INPUT is a stored swish tensor S with the affine ab = (A, B) [n, cin] folded into per-sample
    weights and a border-class bias table, and whose OUTPUT is stored as swish(conv + bias) (act) with the group
    statistics (and optionally the per-channel sums of the stored values, for the SE squeeze)."""
    x2 = None
    if isinstance(x, (tuple, list)):  # channel concat of two dense tensors, read in place by the march kernel
        x, x2 = x
        assert x2.shape[:4] == x.shape[:4] and x2.dtype == torch.bfloat16 and pw.w_march is not None
    n, d, h, w, cin = x.shape
    cin1 = cin
    if x2 is not None:
        cin = cin1 + x2.shape[-1]
    assert x.dtype == torch.bfloat16 and cin == pw.cin and out.shape == (n, d, h, w, pw.cout)
    if ab is None:
        wts = pw.w if pw.taps == 1 else (pw.w_march if pw.w_march is not None else pw.w_slide)
        wstride, table = 0, None
    else:
        wts, wstride, table = _fold_prepare(pw, ab[0], ab[1])
    prof = conv_profile
    if prof is not None:
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
    if pw.taps == 1:
        assert pw.point_ok and chan_sum is None
        call("b21_conv1x1_fwd_fold", ptr(x), _ld(x), ptr(wts), int(wstride != 0), ptr(pw.bias), ptr(table), ptr(out),
             _ld(out), ptr(stats), int(act), n, d * h * w, cin, pw.cout, stream_ptr())
    elif use_input and pw.w_input is not None and ab is None and chan_sum is None and x2 is None and _ld(x) == 8:
        act_code = (2 if fast_input_swish else 1) if act else 0
        call("b21_conv3d_input_fwd", ptr(x), _ld(x), ptr(pw.w_input), ptr(pw.bias), ptr(out), _ld(out), ptr(stats),
             act_code, n, d, h, w, pw.cout, stream_ptr())
    else:
        assert h >= 8 and w >= 8
        name = "b21_conv3d_march_fwd_fold" if pw.w_march is not None else "b21_conv3d_slide_fwd_fold"
        act_code = int(act)
        if act and pw.w_march is not None and cin <= 8 and fast_input_swish:
            act_code = 2  # input conv: 9 MMAs per plane, the epilogue's MUFU work is the bound (fold.cuh: swishf_tanh)
        if x2 is not None:
            call("b21_conv3d_march_fwd_fold2", ptr(x), _ld(x), cin1, ptr(x2), _ld(x2), ptr(wts), wstride, ptr(pw.bias),
                 ptr(table), ptr(out), _ld(out), ptr(stats), ptr(chan_sum), act_code, n, d, h, w, cin, pw.cout,
                 stream_ptr())
        else:
            call(name, ptr(x), _ld(x), ptr(wts), wstride, ptr(pw.bias), ptr(table), ptr(out), _ld(out), ptr(stats),
                 ptr(chan_sum), act_code, n, d, h, w, cin, pw.cout, stream_ptr())
    if prof is not None:
        e1.record()
        kind = "point" if pw.taps == 1 else ("march" if pw.w_march is not None else "slide")
        if use_input and pw.w_input is not None and ab is None and chan_sum is None and x2 is None and _ld(x) == 8:
            kind = "input"
        prof.append((e0, e1, 2.0 * n * d * h * w * pw.cin_true * pw.cout * pw.taps, (kind, cin, pw.cout, pw.taps, d)))
        if kind == "point" and hbm_profile is not None:
            hbm_profile.append((e0, e1, 2.0 * n * d * h * w * (cin + pw.cout), "conv1x1"))
    return out


def affine_pool(x, a_in, b_in, pooled, mode=2):
    """pooled = MaxAvgPool (mode 2) / MaxPool (mode 1) of (A[n][c] * x + B[n][c])."""
    n, d, h, w, c = x.shape
    assert a_in.shape == (n, c) and a_in.stride(0) == b_in.stride(0)
    with _hbm("affine_pool", 2.0 * n * d * h * w * c * (1 + (2 if mode == 2 else 1) / 8)):
        call("b21_affine_pool", ptr(x), _ld(x), ptr(a_in), ptr(b_in), a_in.stride(0), ptr(pooled), _ld(pooled), mode,
             n, d, h, w, c, stream_ptr())
    return pooled


# When set to a list, every conv launch appends (start_event, end_event, algorithmic_flops, shape_key):
# bench.py uses it to time the dominant kernel live with CUDA events on the launching stream.
conv_profile = None
# The plane-marching kernel is the default for the shapes it supports; tools flip this to time the tap kernel.
use_march = True
# im2col first-conv kernel (conv_input.cu) for k = 3 convs with at most 4 real input channels; B21_INPUT_CONV=0 or tests
# flip it back to the plane-marching kernel
use_input = os.environ.get("B21_INPUT_CONV", "1") != "0"
# sliding-window kernel (conv_slide.cu) for the k = 3 shapes whose weights do not fit the march kernel
use_slide = True
# folded-EvoNorm inference path of EquiUnetASSPEvo (csrc/fold.cu); tests flip it to compare both formulations
use_fold = True
# folded EvoNorm on level 3 too (192-channel sliding-window convs with per-sample weights); tests flip it
fold_level3 = True
# level-1 concat of the folded path as two dense tensors (b21_conv3d_march_fwd_fold2); tests flip it
split_concat = True
# single-MUFU swish (tanh.approx) in the epilogue of the Cin = 8 input conv of the folded path; tests flip it
fast_input_swish = True
# replay the inference forward of a window batch from a CUDA graph (networks._B21Net.forward_infer)
use_graphs = True
# plane-marching weight gradient (conv_wgrad_march.cu) for the k = 3 layers with cin <= 96
use_wgrad_march = True
# persistent 1x1 kernel (conv_point.cu) for the shapes it supports
use_point = True
# weight gradients on a side stream, overlapped with the data-gradient chain (autograd.GradStore.on_side)
wgrad_side_stream = True


# ---------------------------------------------------------------------------------------------- norm / SE / pool
GN_RELU, EVO_S0 = 0, 1


CHAN_SLOTS = 16
_chan_slots = {}


def norm_apply(x, stats, gamma, beta, mode, out=None, chan_sum=None, eps=1e-5):
    """GroupNorm(8)+ReLU or EvoNorm-S0 from conv-epilogue statistics; in place when out is None.  chan_sum (fp32 [n, c])
    RECEIVES the per-channel sums of the outputs (the kernel spreads its atomics over CHAN_SLOTS copies)."""
    n, d, h, w, c = x.shape
    if out is None:
        out = x
    slots = None
    if chan_sum is not None:
        key = (n, c, x.device)
        slots = _chan_slots.get(key)
        if slots is None:
            slots = _chan_slots[key] = torch.empty((CHAN_SLOTS, n, c), dtype=torch.float32, device=x.device)
        slots.zero_()
    with _hbm("norm_apply", 4.0 * n * d * h * w * c):
        call("b21_norm_apply", ptr(x), _ld(x), ptr(out), _ld(out), ptr(stats), ptr(gamma), ptr(beta), ptr(slots),
             CHAN_SLOTS, mode, n, d * h * w, c, eps, stream_ptr())
    if slots is not None:
        torch.sum(slots, dim=0, out=chan_sum)
    return out


NORM_GROUP, NORM_INSTANCE, NORM_BATCH_TRAIN, NORM_BATCH_EVAL, NORM_NONE = 0, 1, 2, 3, 4
ACT_CODES = {"none": 0, "relu": 1, "leakyrelu": 2, "elu": 3}


def norm_act(x, kind, act, gamma=None, beta=None, conv_stats=None, running=None, momentum=0.1, out=None, slope=0.01,
             eps=1e-5):
    """The general norm + activation of EquiUnet's factory (networks/factory.py:179-200) after a conv: statistics
    (per channel for instance / batch norm, the conv epilogue's group statistics for GroupNorm) -> per-(n, c) affine ->
    y = act(a * x + b).  running = (running_mean, running_var) for batch norm.  In place when out is None."""
    n, d, h, w, c = x.shape
    nvox = d * h * w
    if out is None:
        out = x
    stats = conv_stats
    if kind in (NORM_INSTANCE, NORM_BATCH_TRAIN):
        stats = torch.empty((n, c, 2), dtype=torch.float64, device=x.device)
        with _hbm("channel_stats", 2.0 * n * nvox * c):
            call("b21_channel_stats", ptr(x), _ld(x), ptr(stats), n, nvox, c, stream_ptr())
    ab = torch.empty((2, n, c), dtype=torch.float32, device=x.device)
    rm, rv = running if running is not None else (None, None)
    call("b21_norm_coeffs", kind, ptr(stats), ptr(gamma), ptr(beta), ptr(rm), ptr(rv), float(momentum), ptr(ab[0]),
         ptr(ab[1]), n, c, nvox, eps, stream_ptr())
    with _hbm("affine_act", 4.0 * n * nvox * c):
        call("b21_affine_act", ptr(x), _ld(x), ptr(out), _ld(out), ptr(ab[0]), ptr(ab[1]), ACT_CODES[act], float(slope),
             n, nvox, c, stream_ptr())
    return out


def se_gate(chan_sum, w1, b1, w2, b2, nvox):
    n, c = chan_sum.shape
    scale = torch.empty_like(chan_sum)
    call("b21_se_gate", ptr(chan_sum), ptr(w1), ptr(b1), ptr(w2), ptr(b2), ptr(scale), n, c, w1.shape[0],
         1.0 / float(nvox), stream_ptr())
    return scale


def scale_pool(x, scale=None, full=None, pooled=None, mode=0):
    """mode 0: full = x*scale; 1: max-pool; 2: [max|avg]-pool (writes 2C channels)."""
    n, d, h, w, c = x.shape
    nb = 2.0 * n * d * h * w * c * (1 + (1 if full is not None else 0) +
                                    ((2 if mode == 2 else 1) / 8 if pooled is not None else 0))
    with _hbm("scale_pool", nb):
        call("b21_scale_pool", ptr(x), _ld(x), ptr(scale), ptr(full), _ld(full) if full is not None else 0,
             ptr(pooled), _ld(pooled) if pooled is not None else 0, mode, n, d, h, w, c, stream_ptr())


def upsample2x(x, out):
    n, d, h, w, c = x.shape
    assert out.shape == (n, 2 * d, 2 * h, 2 * w, c)
    with _hbm("upsample2x", 2.0 * n * d * h * w * c * 9):
        call("b21_upsample2x", ptr(x), _ld(x), ptr(out), _ld(out), n, d, h, w, c, stream_ptr())
    return out


def upsample_f32(x, s, out=None):
    """[N, K, d, h, w] fp32 -> [N, K, s*d, s*h, s*w] (trilinear, align_corners=True)."""
    n, k, d, h, w = x.shape
    if out is None:
        out = torch.empty((n, k, s * d, s * h, s * w), dtype=torch.float32, device=x.device)
    assert out.shape == (n, k, s * d, s * h, s * w) and out.dtype == torch.float32 and out.is_contiguous()
    call("b21_upsample_f32", ptr(x.contiguous()), ptr(out), n * k, d, h, w, s, stream_ptr())
    return out


def head_conv(x, weight, bias, scale=None, out=None, offset=None):
    """1x1 conv to <=4 classes of (x * scale[n] + offset[n]); returns NCDHW fp32 logits.  weight fp32 [K, C]."""
    n, d, h, w, c = x.shape
    k = weight.shape[0]
    if out is None:
        out = torch.empty((n, k, d, h, w), dtype=torch.float32, device=x.device)
    ldso = scale.stride(0) if scale is not None else (offset.stride(0) if offset is not None else c)
    assert offset is None or scale is None or offset.stride(0) == scale.stride(0)
    with _hbm("head_conv", n * d * h * w * (2.0 * c + 4.0 * k)):
        call("b21_head_conv", ptr(x), _ld(x), ptr(scale), ptr(offset), ldso, ptr(weight), ptr(bias), ptr(out), n,
             d * h * w, c, k, stream_ptr())
    return out


# ---------------------------------------------------------------------------------------------- inference wrappers
import ctypes as _C


def _iarr(vals):
    vals = [int(v) for v in vals]
    return (_C.c_int * len(vals))(*vals)


MAX_WINDOWS_PER_CALL = 16  # kMaxWin of csrc/infer.cu (the window list travels as a kernel argument)


def pack_windows(vol, out, origins, perm=(0, 1, 2), flip=(0, 0, 0), vol_index=None):
    """vol [Nv, C, D, H, W] fp32 -> out [B, d, h, w, cpad] bf16 windows of the augmented volume."""
    nv, vc, vd, vh, vw = vol.shape
    b, d, h, w, cpad = out.shape
    assert vol.dtype == torch.float32 and vol.is_contiguous() and out.is_contiguous() and len(origins) == b
    for b0 in range(0, b, MAX_WINDOWS_PER_CALL):
        nb = min(MAX_WINDOWS_PER_CALL, b - b0)
        flat = [c for o in origins[b0:b0 + nb] for c in o]
        vidx = _iarr(vol_index[b0:b0 + nb]) if vol_index is not None else None
        with _hbm("pack_windows", nb * d * h * w * (4.0 * vc + 2.0 * cpad)):
            call("b21_pack_windows", ptr(vol), vc, vd, vh, vw, ptr(out[b0:b0 + nb]), cpad, nb, d, h, w, _iarr(flat), vidx,
                 _iarr(perm), _iarr(flip), stream_ptr())
    return out


def blend_accumulate(logits, acc, profiles, origins, wfloor: float = 0.0):
    """acc [K, AD, AH, AW] += max(outer(profiles), wfloor) * logits [B, K, d, h, w], windows added in order;
    logits None -> count map (K = 1)."""
    k, ad, ah, aw = acc.shape
    pd, ph, pw = profiles
    d, h, w = pd.numel(), ph.numel(), pw.numel()
    nwin = len(origins)
    if logits is not None:
        assert logits.shape == (nwin, k, d, h, w) and logits.is_contiguous() and logits.dtype == torch.float32
    for b0 in range(0, nwin, MAX_WINDOWS_PER_CALL):
        nb = min(MAX_WINDOWS_PER_CALL, nwin - b0)
        flat = [c for o in origins[b0:b0 + nb] for c in o]
        with _hbm("blend_accumulate", nb * k * d * h * w * (12.0 if logits is not None else 8.0)):
            call("b21_blend_accumulate", ptr(logits[b0:b0 + nb]) if logits is not None else None, ptr(acc), ptr(pd),
                 ptr(ph), ptr(pw), nb, k, d, h, w, ad, ah, aw, _iarr(flat), float(wfloor), stream_ptr())


def tta_accumulate(acc, cnt, prob_sum, perm=(0, 1, 2), flip=(0, 0, 0), pad_before=None, apply_sigmoid=True,
                   overwrite=False):
    k, ad, ah, aw = acc.shape
    k2, vd, vh, vw = prob_sum.shape
    assert k == k2 and acc.is_contiguous() and prob_sum.is_contiguous()
    # acc read + prob_sum read-modify-write (write only when overwriting); the count map is read once per voxel
    nb = k * vd * vh * vw * (8.0 if overwrite else 12.0) + (4.0 * vd * vh * vw if cnt is not None else 0.0)
    with _hbm("tta_accumulate", nb):
        call("b21_tta_accumulate", ptr(acc), ptr(cnt), ptr(prob_sum), k, ad, ah, aw,
             _iarr(pad_before) if pad_before is not None else None, vd, vh, vw, _iarr(perm), _iarr(flip),
             int(apply_sigmoid), int(overwrite), stream_ptr())


def labels_finalize(prob_sum, count, thresh=0.5, image=None, want_onehot=True, want_label=True, et_label=4):
    k, vd, vh, vw = prob_sum.shape
    assert k == 3
    nvox = vd * vh * vw
    onehot = torch.empty((3, vd, vh, vw), dtype=torch.uint8, device=prob_sum.device) if want_onehot else None
    label = torch.empty((vd, vh, vw), dtype=torch.uint8, device=prob_sum.device) if want_label else None
    ic = 0
    if image is not None:
        assert image.dtype == torch.float32 and image.is_contiguous() and image.shape[-3:] == (vd, vh, vw)
        ic = image.shape[-4]
    with _hbm("labels_finalize", nvox * (12.0 + 4.0 * ic + (3.0 if want_onehot else 0.0) + (1.0 if want_label else 0.0))):
        call("b21_labels_finalize", ptr(prob_sum), float(count), float(thresh), ptr(image), ic, ptr(onehot), ptr(label),
             nvox, et_label, stream_ptr())
    return onehot, label


def mask_background(label, image):
    """In place: zero the uint8 label planes [..., D, H, W] where every channel of image [C, D, H, W] is 0."""
    nvox = image[0].numel()
    assert label.dtype == torch.uint8 and label.is_contiguous() and label.numel() % nvox == 0
    assert image.dtype == torch.float32 and image.is_contiguous()
    call("b21_mask_background", ptr(label), label.numel() // nvox, ptr(image), image.shape[0], nvox, stream_ptr())
    return label


# ---------------------------------------------------------------------------------------------- training step
def _ld4(t: torch.Tensor) -> int:
    return _ld(t)


def conv3d_wgrad(x, dz, dw, dil: int = 1):
    """dw (fp32 [cout, cin, k, k, k], contiguous) += weight gradient of conv3d(x) given dz = d(out)."""
    n, d, h, w, cin = x.shape
    cout = dz.shape[-1]
    k = dw.shape[2]
    assert dw.dtype == torch.float32 and dw.is_contiguous() and dw.shape[0] == cout and dw.shape[1] <= cin
    if use_wgrad_march and k == 3 and dil == 1 and h >= 8 and w >= 8 and \
            _lib.load().b21_conv_wgrad_march_supported(cin, cout):
        prof = conv_profile
        if prof is not None:
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
        call("b21_conv3d_wgrad_march", ptr(x), _ld(x), ptr(dz), _ld(dz), ptr(dw), n, d, h, w, cin, dw.shape[1], cout,
             stream_ptr())
        if prof is not None:
            e1.record()
            prof.append((e0, e1, 2.0 * n * d * h * w * dw.shape[1] * cout * 27, ("wgrad_march", cin, cout, 27, d)))
        return dw
    if dw.shape[1] != cin:  # input channels were zero-padded for the forward pass (first layer): use a padded scratch
        tmp = torch.zeros((cout, cin) + tuple(dw.shape[2:]), dtype=torch.float32, device=dw.device)
        call("b21_conv3d_wgrad", ptr(x), _ld(x), ptr(dz), _ld(dz), ptr(tmp), n, d, h, w, cin, cout, k ** 3, dil,
             stream_ptr())
        dw += tmp[:, :dw.shape[1]]
        return dw
    prof = conv_profile
    if prof is not None:
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
    call("b21_conv3d_wgrad", ptr(x), _ld(x), ptr(dz), _ld(dz), ptr(dw), n, d, h, w, cin, cout, k ** 3, dil, stream_ptr())
    if prof is not None:
        e1.record()
        prof.append((e0, e1, 2.0 * n * d * h * w * cin * cout * k ** 3, ("wgrad", cin, cout, k ** 3, d)))
    return dw


def norm_bwd_workspace(n, c, device):
    return torch.empty((int(_lib.load().b21_norm_bwd_workspace_bytes(n, c)),), dtype=torch.uint8, device=device)


def norm_bwd(dy, z, dz, stats, gamma, beta, dgamma, dbeta, mode, colsum=None, se=None, workspace=None, eps=1e-5):
    """Backward of norm_apply (+ optional squeeze-excite gate).  se = dict(scale, mean, w1, b1, w2, b2, dw1, db1, dw2,
    db2) with fp32 contiguous tensors.  Accumulates dgamma/dbeta/colsum/SE grads; writes dz (may alias dy)."""
    n, d, h, w, c = z.shape
    if workspace is None:
        workspace = norm_bwd_workspace(n, c, z.device)
    s = se or {}
    call("b21_norm_bwd", ptr(dy), _ld(dy), ptr(z), _ld(z), ptr(dz), _ld(dz), ptr(stats), ptr(gamma), ptr(beta),
         ptr(dgamma), ptr(dbeta), ptr(colsum), ptr(s.get("scale")), ptr(s.get("mean")), ptr(s.get("w1")),
         ptr(s.get("b1")), ptr(s.get("w2")), ptr(s.get("b2")), ptr(s.get("dw1")), ptr(s.get("db1")), ptr(s.get("dw2")),
         ptr(s.get("db2")), s["w1"].shape[0] if se else 0, ptr(workspace), workspace.numel(), mode, n, d * h * w, c, eps,
         stream_ptr())
    return dz


def pool_bwd(y, dpool, dy, mode, add=None):
    n, d, h, w, c = y.shape
    call("b21_pool_bwd", ptr(y), _ld(y), ptr(dpool), _ld(dpool), ptr(add), _ld(add) if add is not None else 0, ptr(dy),
         _ld(dy), mode, n, d, h, w, c, stream_ptr())
    return dy


def upsample2x_bwd(dy, dx):
    n, d, h, w, c = dx.shape
    assert dy.shape == (n, 2 * d, 2 * h, 2 * w, c)
    call("b21_upsample2x_bwd", ptr(dy), _ld(dy), ptr(dx), _ld(dx), n, d, h, w, c, stream_ptr())
    return dx


def upsample_f32_bwd(dy, s):
    n, k, do, ho, wo = dy.shape
    d, h, w = do // s, ho // s, wo // s
    dx = torch.empty((n, k, d, h, w), dtype=torch.float32, device=dy.device)
    call("b21_upsample_f32_bwd", ptr(dy.contiguous()), ptr(dx), n * k, d, h, w, s, stream_ptr())
    return dx


def head_conv_bwd(x, weight, dl, dx, scale=None, accumulate=False):
    """Returns (dws [n, k, c], db [k]); writes/accumulates dx (bf16)."""
    n, d, h, w, c = x.shape
    k = weight.shape[0]
    # CHAN_SLOTS copies of both tables (block b adds to copy b % slots), summed below
    dws_t = torch.zeros((CHAN_SLOTS, n, k, c), dtype=torch.float32, device=x.device)
    db_t = torch.zeros((CHAN_SLOTS, k), dtype=torch.float32, device=x.device)
    call("b21_head_conv_bwd", ptr(x), _ld(x), ptr(scale), ptr(weight), ptr(dl.contiguous()), ptr(dx), _ld(dx),
         int(accumulate), ptr(dws_t), ptr(db_t), CHAN_SLOTS, n, d * h * w, c, k, stream_ptr())
    return dws_t.sum(0), db_t.sum(0)


def add_inplace(dst, src):
    n, d, h, w, c = dst.shape
    assert src.shape == dst.shape
    call("b21_add_inplace", ptr(dst), _ld(dst), ptr(src), _ld(src), n * d * h * w, c, stream_ptr())
    return dst


def dice_fwd(logits, target, loss, jaccard=False, weight=1.0, smooth_nr=1e-5, smooth_dr=1e-5):
    """loss[0] += weight * DiceLoss(logits, target); returns the coefficient table for dice_bwd."""
    n, k = logits.shape[:2]
    nvox = logits[0, 0].numel()
    sums = torch.empty((k, 3), dtype=torch.float64, device=logits.device)
    coef = torch.empty((k, 2), dtype=torch.float32, device=logits.device)
    call("b21_dice_fwd", ptr(logits), ptr(target), ptr(sums), ptr(loss), ptr(coef), n, k, nvox, int(jaccard),
         smooth_nr, smooth_dr, weight, stream_ptr())
    return coef


def ce_fwd(logits, target, loss, weight=1.0):
    """loss[0] += weight * CrossEntropy(logits, argmax_k target) (mean over batch and voxels)."""
    n, k = logits.shape[:2]
    scratch = torch.empty((1,), dtype=torch.float64, device=logits.device)
    call("b21_ce_fwd", ptr(logits), ptr(target), ptr(scratch), ptr(loss), n, k, logits[0, 0].numel(), weight,
         stream_ptr())


def dice_bwd(logits, target, coef, gout=None, gscale=1.0, ce_weight=0.0):
    n, k = logits.shape[:2]
    dl = torch.empty_like(logits)
    call("b21_dice_bwd", ptr(logits), ptr(target), ptr(coef), ptr(gout), gscale, float(ce_weight), ptr(dl), n, k,
         logits[0, 0].numel(), stream_ptr())
    return dl
