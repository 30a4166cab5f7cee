"""Training path of the drop-in networks: forward that keeps what the backward needs, and a hand-scheduled backward
on the libb21 kernels (learning/engine.py:88-130 is ``zero_grad -> model(img) -> criterion -> backward -> step``).

``network_forward_train(net, x)`` returns fp32 logits ``(out, [deeps])`` attached to ONE autograd node; when the
criterion's gradient reaches that node, ``_backward_v2`` walks the network in reverse with
  * data gradients   = the forward conv kernels on the transposed+mirrored packed weights,
  * weight gradients = the split-K MN-major tcgen05 kernel (conv_wgrad.cu),
  * EvoNorm-S0 (+ squeeze-excite) backward as one reduction pass and one apply pass (train.cu),
  * pool / trilinear / head adjoints (train.cu),
and ACCUMULATES parameter gradients (fp32) into one flat buffer laid out in backward-completion order, of which every
``param.grad`` is a view — the data-parallel wrapper all-reduces that buffer bucket by bucket while the rest of the
backward is still running (brats21_b200/parallel.py).  Nothing here falls back to eager PyTorch modules.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional

import torch

from . import ops

EVO = ops.EVO_S0


# ================================================================================================ gradient store
class GradStore:
    """Flat fp32 gradient buffer; ``views[name]`` has the parameter's shape.  Order = backward completion order."""

    def __init__(self, net, order: List[str]):
        params = dict(net.named_parameters())
        self.names = [n for n in order if n in params]
        missing = [n for n, p in params.items() if n not in self.names and p.requires_grad and not n.endswith(".v")]
        if missing:
            raise RuntimeError(f"parameters without a place in the backward schedule: {missing}")
        self.params = [params[n] for n in self.names]
        # every view starts on a 16-byte boundary (the kernels that accumulate into the views use 16 B accesses); the
        # padding elements stay zero and ride along in the all-reduce
        pad4 = lambda k: (k + 3) // 4 * 4  # noqa: E731
        total = sum(pad4(p.numel()) for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros((total,), dtype=torch.float32, device=dev)
        self.views: Dict[str, torch.Tensor] = {}
        self.offsets: Dict[str, int] = {}
        off = 0
        for n, p in zip(self.names, self.params):
            self.views[n] = self.flat[off:off + p.numel()].view(p.shape)
            self.offsets[n] = off
            off += pad4(p.numel())
        # data-parallel hooks (brats21_b200.parallel.BucketReducer): start of a backward, "gradients in
        # flat[0:offset] are final", end of the backward
        self.on_begin: Optional[Callable[[], None]] = None
        self.on_ready: Optional[Callable[..., None]] = None
        self.on_finish: Optional[Callable[[], None]] = None
        # Weight gradients run on a SIDE stream: dW of a layer only feeds the optimizer, while the data gradient and
        # the HBM-bound norm / pool adjoints of the next layers are on the critical path.  The tensor-bound wgrad
        # kernels (one persistent CTA per SM, ~200 KB of shared memory each) co-reside with the register-only
        # element-wise kernels, so the two streams fill each other's idle pipes.
        self.side: Optional[torch.cuda.Stream] = None
        self._side_dirty = False

    def begin(self):
        """Bind .grad views; zero the buffer unless the caller is accumulating into existing gradients."""
        fresh = all(p.grad is None or p.grad.data_ptr() != v.data_ptr() for p, v in zip(self.params, self.views.values()))
        if fresh:
            self.flat.zero_()
            for p, n in zip(self.params, self.names):
                p.grad = self.views[n]
        if self.on_begin is not None:
            self.on_begin()

    def on_side(self, fn: Callable[[], None]):
        """Run ``fn`` (kernel launches that only write parameter gradients) on the side stream, ordered after
        everything enqueued so far on the current stream."""
        if not (ops.wgrad_side_stream and self.flat.is_cuda) or ops.conv_profile is not None:
            fn()
            return
        if self.side is None:
            self.side = torch.cuda.Stream(device=self.flat.device)
        ev = torch.cuda.Event()
        ev.record()
        with torch.cuda.stream(self.side):
            self.side.wait_event(ev)
            fn()
        self._side_dirty = True

    def side_fence(self):
        """Event after the work enqueued on the side stream so far (None when nothing is pending there)."""
        if not self._side_dirty:
            return None
        ev = torch.cuda.Event()
        ev.record(self.side)
        return ev

    def join_side(self):
        """The current stream waits for everything enqueued on the side stream."""
        if self._side_dirty:
            torch.cuda.current_stream().wait_stream(self.side)
            self._side_dirty = False

    def _side_event(self):
        if not self._side_dirty:
            return ()
        ev = torch.cuda.Event()
        ev.record(self.side)
        return (ev,)

    def finish(self):
        if self._side_dirty:
            torch.cuda.current_stream().wait_stream(self.side)
            self._side_dirty = False
        if self.on_finish is not None:
            self.on_finish()

    def ready(self, name: str):
        """Gradients up to and including ``name`` (backward-completion order) are final once the work enqueued so far
        on the current AND the side stream has run."""
        if self.on_ready is not None:
            self.on_ready(self.offsets[name] + self.views[name].numel(), self._side_event())


# ================================================================================================ V2
_V2_BLOCKS_BWD = ["decoder1", "decoder2", "decoder3", "encoder4", "encoder3", "encoder2", "encoder1"]


def v2_grad_order(net) -> List[str]:
    """Parameter names in the order their gradients become final during _backward_v2."""
    def block(b):
        s = b + ".conv_conv_se."
        return [s + "6.fc.0.weight", s + "6.fc.0.bias", s + "6.fc.2.weight", s + "6.fc.2.bias", s + "4.gamma", s + "4.beta",
                s + "3.weight", s + "3.bias", s + "1.gamma", s + "1.beta", s + "0.weight", s + "0.bias"]

    def convevo(c):
        return [c + ".evo.gamma", c + ".evo.beta", c + ".conv.weight", c + ".conv.bias"]

    order: List[str] = []
    if net.deep_supervision:
        order += ["deep3.0.weight", "deep3.0.bias", "deep2.0.weight", "deep2.0.bias"]
    order += ["out_conv.weight", "out_conv.bias"]
    order += block("decoder1") + convevo("upconv1") + convevo("bridge1")
    order += block("decoder2") + convevo("upconv2") + convevo("bridge2")
    order += block("decoder3") + convevo("upconv3") + convevo("bridge3")
    order += convevo("aspp.conv_k1")
    for i in range(4):
        order += [f"aspp.convs.{i}.weight", f"aspp.convs.{i}.bias"]
    order += block("encoder4") + block("encoder3") + block("encoder2") + block("encoder1")
    return order


def _pack_train_v2(net):
    """Transposed + mirrored packings for the data gradients (every conv except the very first)."""
    pk = net._packed
    for b in net._BLOCKS:
        seq = net.get_submodule(b).conv_conv_se
        if b != "encoder1":
            pk[b + ".c0.T"] = ops.PackedConv(seq[0].weight, None, transpose_flip=True)
        pk[b + ".c1.T"] = ops.PackedConv(seq[3].weight, None, transpose_flip=True)
    for name in net._CONVEVO:
        pk[name + ".T"] = ops.PackedConv(net.get_submodule(name).conv.weight, None, transpose_flip=True)
    for i, conv in enumerate(net.aspp.convs):
        pk[f"aspp.convs.{i}.T"] = ops.PackedConv(conv.weight, None, transpose_flip=True)
    pk["__train__"] = True
    net._sync_pack_table()


def _v2_forward_train(net, x8: torch.Tensor, want_deep: bool):
    net._ensure_packed()
    pk = net._packed
    if "__train__" not in pk:
        _pack_train_v2(net)
    n, d, h, w, _ = x8.shape
    f = net.features
    ws = net._ws.setdefault(("v2train", n, d, h, w), {})
    B = lambda name, s, c: net._buf(ws, name, (n, d // s, h // s, w // s, c))  # noqa: E731
    S = lambda name: net._buf(ws, name, (ops._lib.STAT_SLOTS, n, 8, 2), torch.float64)  # noqa: E731
    tape: Dict[str, dict] = {}

    def block(name, x, s, c):
        z0, a0, z1, y = B(name + ".z0", s, c), B(name + ".a0", s, c), B(name + ".z1", s, c), B(name + ".y", s, c)
        st0, st1 = S(name + ".st0"), S(name + ".st1")
        csum = net._buf(ws, name + ".csum", (n, c), torch.float32)
        ops.conv3d(x, pk[name + ".c0"], out=z0, stats=st0)
        ops.norm_apply(z0, st0, pk[name + ".e0.g"], pk[name + ".e0.b"], EVO, out=a0)
        ops.conv3d(a0, pk[name + ".c1"], out=z1, stats=st1)
        csum.zero_()
        ops.norm_apply(z1, st1, pk[name + ".e1.g"], pk[name + ".e1.b"], EVO, out=y, chan_sum=csum)
        nvox = (d // s) * (h // s) * (w // s)
        scale = ops.se_gate(csum, pk[name + ".se.w1"], pk[name + ".se.b1"], pk[name + ".se.w2"], pk[name + ".se.b2"], nvox)
        tape[name] = dict(x=x, z0=z0, a0=a0, z1=z1, y=y, st0=st0, st1=st1, scale=scale, mean=csum / float(nvox))
        return y, scale

    def convevo(name, x, out, s, c):
        z, st = B(name + ".z", s, c), S(name + ".st")
        ops.conv3d(x, pk[name], out=z, stats=st)
        ops.norm_apply(z, st, pk[name + ".g"], pk[name + ".b"], EVO, out=out)
        tape[name] = dict(x=x, z=z, st=st)
        return out

    p1, p2, p3 = B("p1", 2, 2 * f[0]), B("p2", 4, 2 * f[1]), B("p3", 8, 2 * f[2])
    cat1, cat2, cat3 = B("cat1", 1, f[0]), B("cat2", 2, f[1]), B("cat3", 4, f[2])
    y1, s = block("encoder1", x8, 1, f[0])
    ops.scale_pool(y1, s, full=y1, pooled=p1, mode=2)
    y2, s = block("encoder2", p1, 2, f[1])
    ops.scale_pool(y2, s, full=y2, pooled=p2, mode=2)
    y3, s = block("encoder3", p2, 4, f[2])
    ops.scale_pool(y3, s, full=y3, pooled=p3, mode=2)
    y4, s = block("encoder4", p3, 8, f[3])
    ops.scale_pool(y4, s, full=y4, mode=0)

    acat = B("asppcat", 8, f[3])
    q = f[3] // 4
    for i, dil in enumerate(net.aspp.dilations):
        ops.conv3d(y4, pk[f"aspp.convs.{i}"], out=acat[..., i * q:(i + 1) * q], dil=dil)
    assp = convevo("aspp.conv_k1", acat, B("assp", 8, f[3]), 8, f[3])

    convevo("bridge1", y1, cat1[..., :f[0] // 2], 1, f[0] // 2)
    convevo("bridge2", y2, cat2[..., :f[1] // 2], 2, f[1] // 2)
    convevo("bridge3", y3, cat3[..., :f[2] // 2], 4, f[2] // 2)

    # Deep-supervision heads (1x1 conv to 3 classes + fp32 trilinear x4 / x2): HBM-bound launches that only the loss
    # needs.  They run on the side stream under the tensor-bound decoder convs; outputs are allocated on the current
    # stream, which waits for the side stream before the forward returns.
    gs = net.grad_store()
    deep_on = want_deep and net.deep_supervision
    deeps: List[torch.Tensor] = []

    def deep_head(name, src, scale):
        nk, (sd, sh, sw) = net.num_classes, src.shape[1:4]
        up = torch.empty((n, nk, sd * scale, sh * scale, sw * scale), dtype=torch.float32, device=src.device)

        def run():
            # The low-resolution logits are a temporary of the SIDE stream and must be allocated under it: a block taken
            # from the current stream's pool goes back to that pool as soon as the last reference dies — while the side
            # stream may still be reading it — and the next small allocation of the main stream (SE gate, channel means)
            # lands on top of it.  (Seen as a deep-supervision loss that was off by 0.1 % in about half of the runs of the
            # two-process data-parallel test, profiles/r02z_grad_noise.md.)
            low = torch.empty((n, nk, sd, sh, sw), dtype=torch.float32, device=src.device)
            ops.upsample_f32(ops.head_conv(src, pk[name + ".w"], pk[name + ".bias"], out=low), scale, out=up)

        gs.on_side(run)
        deeps.append(up)

    u = convevo("upconv3", assp, B("uc3", 8, f[3] // 4), 8, f[3] // 4)
    ops.upsample2x(u, cat3[..., f[2] // 2:])
    yd3, s = block("decoder3", cat3, 4, f[2])
    ops.scale_pool(yd3, s, full=yd3, mode=0)
    if deep_on:
        deep_head("deep3.0", yd3, 4)
    u = convevo("upconv2", yd3, B("uc2", 4, f[2] // 4), 4, f[2] // 4)
    ops.upsample2x(u, cat2[..., f[1] // 2:])
    yd2, s = block("decoder2", cat2, 2, f[1])
    ops.scale_pool(yd2, s, full=yd2, mode=0)
    if deep_on:
        deep_head("deep2.0", yd2, 2)
    u = convevo("upconv1", yd2, B("uc1", 2, f[1] // 4), 2, f[1] // 4)
    ops.upsample2x(u, cat1[..., f[0] // 2:])
    a1, s1 = block("decoder1", cat1, 1, f[0])
    out = ops.head_conv(a1, pk["out_conv.w"], pk["out_conv.bias"], scale=s1)
    gs.join_side()
    tape["__meta__"] = dict(shape=(n, d, h, w), ws=ws, y=(y1, y2, y3, y4), yd=(yd3, yd2, a1), s1=s1, acat=acat, assp=assp,
                            cats=(cat1, cat2, cat3), want_deep=bool(deeps))
    return out, deeps, tape


def _backward_v2(net, tape, dout: Optional[torch.Tensor], ddeeps: List[Optional[torch.Tensor]], gs: GradStore):
    pk, G = net._packed, gs.views
    meta = tape["__meta__"]
    n, d, h, w = meta["shape"]
    ws, f = meta["ws"], net.features
    B = lambda name, s, c: net._buf(ws, "g." + name, (n, d // s, h // s, w // s, c))  # noqa: E731
    y1, y2, y3, y4 = meta["y"]
    yd3, yd2, a1 = meta["yd"]
    cat1, cat2, cat3 = meta["cats"]
    nbw = ops.norm_bwd_workspace(n, f[3], a1.device)

    def evo_bwd(gname, bname, dy, z, st, dz, colsum, se=None):
        ops.norm_bwd(dy, z, dz, st, pk[gname], pk[bname], G_flat(gname), G_flat(bname), EVO, colsum=colsum, se=se,
                     workspace=nbw)

    # parameter-name lookup for the packed fp32 vectors (gamma/beta keys of the packed dict -> state_dict names)
    def G_flat(packed_key):
        return G[_V2_KEYMAP(net)[packed_key]].view(-1)

    def block_bwd(name, dy, dx_out, s, c):
        t = tape[name]
        pre = name + ".conv_conv_se."
        se = dict(scale=t["scale"], mean=t["mean"], w1=pk[name + ".se.w1"], b1=pk[name + ".se.b1"], w2=pk[name + ".se.w2"],
                  b2=pk[name + ".se.b2"], dw1=G[pre + "6.fc.0.weight"], db1=G[pre + "6.fc.0.bias"],
                  dw2=G[pre + "6.fc.2.weight"], db2=G[pre + "6.fc.2.bias"])
        # Order on the two streams: the data gradient (tensor-bound, on the critical path) is enqueued BEFORE the weight
        # gradient of the same layer is released to the side stream.  Both are persistent kernels that own an SM's
        # shared memory, so they can only run one after the other; released together (as round 1 did) the data gradient
        # waited for the weight gradient and the HBM-bound norm adjoint that follows ran alone.  Now the weight gradient
        # runs under the norm adjoint of the next layer (timeline: profiles/r02l_timeline_v2_train.md).
        evo_bwd(name + ".e1.g", name + ".e1.b", dy, t["z1"], t["st1"], dy, G[pre + "3.bias"], se=se)   # dy <- dz1
        da0 = B(name + ".da0", s, c)
        ops.conv3d(dy, pk[name + ".c1.T"], out=da0)
        gs.on_side(lambda: ops.conv3d_wgrad(t["a0"], dy, G[pre + "3.weight"]))
        evo_bwd(name + ".e0.g", name + ".e0.b", da0, t["z0"], t["st0"], da0, G[pre + "0.bias"])       # da0 <- dz0
        if dx_out is not None:
            ops.conv3d(da0, pk[name + ".c0.T"], out=dx_out)
        gs.on_side(lambda: ops.conv3d_wgrad(t["x"], da0, G[pre + "0.weight"]))
        gs.ready(pre + "0.bias")

    def convevo_bwd(name, dy, dz, dx_out):
        t = tape[name]
        evo_bwd(name + ".g", name + ".b", dy, t["z"], t["st"], dz, G[name + ".conv.bias"])
        if dx_out is not None:
            ops.conv3d(dz, pk[name + ".T"], out=dx_out)
        gs.on_side(lambda: ops.conv3d_wgrad(t["x"], dz, G[name + ".conv.weight"]))
        gs.ready(name + ".conv.bias")

    def head_bwd(pname, x, dl, dx, scale_fold=None):
        wkey = pname + ".w"
        dws, db = ops.head_conv_bwd(x, pk[wkey], dl, dx, scale=None, accumulate=False)
        if scale_fold is not None:
            dws = dws * scale_fold[:, None, :]
        G[pname + ".weight"].view(dws.shape[1], dws.shape[2]).add_(dws.sum(0))
        G[pname + ".bias"].add_(db)
        gs.ready(pname + ".bias")

    # ---- heads
    g_yd3, g_yd2, g_a1 = B("yd3", 4, f[2]), B("yd2", 2, f[1]), B("a1", 1, f[0])
    if meta["want_deep"]:
        dl3 = ddeeps[0] if ddeeps[0] is not None else None
        dl2 = ddeeps[1] if ddeeps[1] is not None else None
        # the adjoints of the deep heads are first needed at decoder2 / decoder3: side stream, under decoder1's backward
        if dl3 is not None:
            gs.on_side(lambda: head_bwd("deep3.0", yd3, ops.upsample_f32_bwd(dl3.float(), 4), g_yd3))
        else:
            g_yd3.zero_()
        if dl2 is not None:
            gs.on_side(lambda: head_bwd("deep2.0", yd2, ops.upsample_f32_bwd(dl2.float(), 2), g_yd2))
        else:
            g_yd2.zero_()
    else:
        g_yd3.zero_()
        g_yd2.zero_()
    deep_fence = gs.side_fence()
    if dout is None:
        dout = torch.zeros((n, net.num_classes, d, h, w), dtype=torch.float32, device=a1.device)
    head_bwd("out_conv", a1, dout.float(), g_a1, scale_fold=meta["s1"])

    # ---- decoder 1 and its inputs
    g_cat1 = B("cat1", 1, f[0])
    block_bwd("decoder1", g_a1, g_cat1, 1, f[0])
    g_uc1, tmp2 = B("uc1", 2, f[1] // 4), B("tmp2", 2, f[1])
    ops.upsample2x_bwd(g_cat1[..., f[0] // 2:], g_uc1)
    convevo_bwd("upconv1", g_uc1, g_uc1, tmp2)
    if deep_fence is not None:
        torch.cuda.current_stream().wait_event(deep_fence)  # g_yd2 / g_yd3 hold the deep heads' gradients from here on
    ops.add_inplace(g_yd2, tmp2)
    g_b1, g_y1 = B("b1", 1, f[0] // 2), B("y1", 1, f[0])
    convevo_bwd("bridge1", g_cat1[..., :f[0] // 2], g_b1, g_y1)

    # ---- decoder 2
    g_cat2 = B("cat2", 2, f[1])
    block_bwd("decoder2", g_yd2, g_cat2, 2, f[1])
    g_uc2, tmp3 = B("uc2", 4, f[2] // 4), B("tmp3", 4, f[2])
    ops.upsample2x_bwd(g_cat2[..., f[1] // 2:], g_uc2)
    convevo_bwd("upconv2", g_uc2, g_uc2, tmp3)
    ops.add_inplace(g_yd3, tmp3)
    g_b2, g_y2 = B("b2", 2, f[1] // 2), B("y2", 2, f[1])
    convevo_bwd("bridge2", g_cat2[..., :f[1] // 2], g_b2, g_y2)

    # ---- decoder 3
    g_cat3 = B("cat3", 4, f[2])
    block_bwd("decoder3", g_yd3, g_cat3, 4, f[2])
    g_uc3, g_assp = B("uc3", 8, f[3] // 4), B("assp", 8, f[3])
    ops.upsample2x_bwd(g_cat3[..., f[2] // 2:], g_uc3)
    convevo_bwd("upconv3", g_uc3, g_uc3, g_assp)
    g_b3, g_y3 = B("b3", 4, f[2] // 2), B("y3", 4, f[2])
    convevo_bwd("bridge3", g_cat3[..., :f[2] // 2], g_b3, g_y3)

    # ---- ASPP
    g_acat, g_y4, tmp4 = B("acat", 8, f[3]), B("y4", 8, f[3]), B("tmp4", 8, f[3])
    convevo_bwd("aspp.conv_k1", g_assp, g_assp, g_acat)
    q = f[3] // 4
    for i, dil in enumerate(net.aspp.dilations):
        dz = g_acat[..., i * q:(i + 1) * q]
        G[f"aspp.convs.{i}.bias"].add_(dz.float().sum(dim=(0, 1, 2, 3)))
        ops.conv3d(dz, pk[f"aspp.convs.{i}.T"], out=g_y4 if i == 0 else tmp4, dil=dil)
        gs.on_side(lambda dz=dz, i=i, dil=dil: ops.conv3d_wgrad(y4, dz, G[f"aspp.convs.{i}.weight"], dil=dil))
        if i > 0:
            ops.add_inplace(g_y4, tmp4)
        gs.ready(f"aspp.convs.{i}.bias")

    # ---- encoders
    g_p3 = B("p3", 8, 2 * f[2])
    block_bwd("encoder4", g_y4, g_p3, 8, f[3])
    ops.pool_bwd(y3, g_p3, g_y3, 2, add=g_y3)
    g_p2 = B("p2", 4, 2 * f[1])
    block_bwd("encoder3", g_y3, g_p2, 4, f[2])
    ops.pool_bwd(y2, g_p2, g_y2, 2, add=g_y2)
    g_p1 = B("p1", 2, 2 * f[0])
    block_bwd("encoder2", g_y2, g_p1, 2, f[1])
    ops.pool_bwd(y1, g_p1, g_y1, 2, add=g_y1)
    block_bwd("encoder1", g_y1, None, 1, f[0])


_KEYMAPS: Dict[int, Dict[str, str]] = {}


def _V2_KEYMAP(net) -> Dict[str, str]:
    km = _KEYMAPS.get(id(net))
    if km is None:
        km = {}
        for b in net._BLOCKS:
            pre = b + ".conv_conv_se."
            km[b + ".e0.g"], km[b + ".e0.b"] = pre + "1.gamma", pre + "1.beta"
            km[b + ".e1.g"], km[b + ".e1.b"] = pre + "4.gamma", pre + "4.beta"
        for c in net._CONVEVO:
            km[c + ".g"], km[c + ".b"] = c + ".evo.gamma", c + ".evo.beta"
        _KEYMAPS[id(net)] = km
    return km


# ================================================================================================ autograd node
class _NetFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, net, x, *params):
        x8 = net.pack_input(x)
        out, deeps, tape = net._forward_train(x8, True)
        ctx.net, ctx.tape, ctx.ndeep = net, tape, len(deeps)
        return (out, *deeps)

    @staticmethod
    def backward(ctx, dout, *ddeeps):
        net = ctx.net
        gs = net.grad_store()
        gs.begin()
        with torch.no_grad():
            net._backward_train(ctx.tape, dout, list(ddeeps), gs)
            gs.finish()
        ctx.tape = None
        return (None, None) + (None,) * len(list(net.parameters()))


def network_forward_train(net, x: torch.Tensor):
    """forward of ``net`` in training mode; returns (out, [deeps]) like the reference (deep_supervision) or out."""
    params = list(net.parameters())
    res = _NetFn.apply(net, x, *params)
    out, deeps = res[0], list(res[1:])
    if net.deep_supervision:
        return out, deeps
    return out


# ================================================================================================ V1
GN = ops.GN_RELU
_V1_HEADS = (("deep_bottom.0", 8), ("deep_bottom2.0", 8), ("deep3.0", 4), ("deep2.0", 2))


def v1_grad_order(net) -> List[str]:
    def cbr(c):
        return [c + ".bn.weight", c + ".bn.bias", c + ".conv.weight"]

    def ublock(b):
        return cbr(b + ".ConvBnRelu2") + cbr(b + ".ConvBnRelu1")

    order = ["outconv.weight", "outconv.bias"]
    order += ublock("decoder1")
    if net.deep_supervision:
        order += ["deep2.0.weight", "deep2.0.bias"]
    order += ublock("decoder2")
    if net.deep_supervision:
        order += ["deep3.0.weight", "deep3.0.bias"]
    order += ublock("decoder3")
    if net.deep_supervision:
        order += ["deep_bottom2.0.weight", "deep_bottom2.0.bias"]
    order += cbr("bottom_2")
    if net.deep_supervision:
        order += ["deep_bottom.0.weight", "deep_bottom.0.bias"]
    order += ublock("bottom") + ublock("encoder4") + ublock("encoder3") + ublock("encoder2") + ublock("encoder1")
    return order


def _v1_forward_train(net, x8: torch.Tensor, want_deep: bool):
    net._ensure_packed()
    pk = net._packed
    if "__train__" not in pk:
        for name in net._CBR:
            if name != "encoder1.ConvBnRelu1":
                pk[name + ".T"] = ops.PackedConv(net.get_submodule(name).conv.weight, None, transpose_flip=True)
        pk["__train__"] = True
        net._sync_pack_table()
    n, d, h, w, _ = x8.shape
    f = net.features
    ws = net._ws.setdefault(("v1train", n, d, h, w), {})
    B = lambda name, s, c: net._buf(ws, name, (n, d // s, h // s, w // s, c))  # noqa: E731
    tape: Dict[str, dict] = {}

    def cbr(name, x, out, s, dil=1):
        c = out.shape[-1]
        z = B(name + ".z", s, c)
        st = net._buf(ws, name + ".st", (ops._lib.STAT_SLOTS, n, 8, 2), torch.float64)
        ops.conv3d(x, pk[name], out=z, stats=st, dil=dil)
        ops.norm_apply(z, st, pk[name + ".g"], pk[name + ".b"], GN, out=out)
        tape[name] = dict(x=x, z=z, st=st, dil=dil)
        return out

    cat1, cat2, cat3, cat4 = B("cat1", 1, 2 * f[0]), B("cat2", 2, 2 * f[1]), B("cat3", 4, 2 * f[2]), B("cat4", 8, 2 * f[3])
    p1, p2, p3 = B("p1", 2, f[0]), B("p2", 4, f[1]), B("p3", 8, f[2])
    down1 = cbr("encoder1.ConvBnRelu2", cbr("encoder1.ConvBnRelu1", x8, B("e1", 1, f[0]), 1), cat1[..., :f[0]], 1)
    ops.scale_pool(down1, pooled=p1, mode=1)
    down2 = cbr("encoder2.ConvBnRelu2", cbr("encoder2.ConvBnRelu1", p1, B("e2", 2, f[1]), 2), cat2[..., :f[1]], 2)
    ops.scale_pool(down2, pooled=p2, mode=1)
    down3 = cbr("encoder3.ConvBnRelu2", cbr("encoder3.ConvBnRelu1", p2, B("e3", 4, f[2]), 4), cat3[..., :f[2]], 4)
    ops.scale_pool(down3, pooled=p3, mode=1)
    down4 = cbr("encoder4.ConvBnRelu2", cbr("encoder4.ConvBnRelu1", p3, B("e4", 8, f[3]), 8), cat4[..., :f[3]], 8)
    bottom = cbr("bottom.ConvBnRelu2", cbr("bottom.ConvBnRelu1", down4, B("bt", 8, f[3]), 8, dil=2), cat4[..., f[3]:], 8,
                 dil=2)
    b2 = cbr("bottom_2", cat4, B("b2", 8, f[2]), 8)
    ops.upsample2x(b2, cat3[..., f[2]:])
    u3 = cbr("decoder3.ConvBnRelu2", cbr("decoder3.ConvBnRelu1", cat3, B("d3", 4, f[2]), 4), B("u3", 4, f[1]), 4)
    ops.upsample2x(u3, cat2[..., f[1]:])
    u2 = cbr("decoder2.ConvBnRelu2", cbr("decoder2.ConvBnRelu1", cat2, B("d2", 2, f[1]), 2), B("u2", 2, f[0]), 2)
    ops.upsample2x(u2, cat1[..., f[0]:])
    u1 = cbr("decoder1.ConvBnRelu2", cbr("decoder1.ConvBnRelu1", cat1, B("d1", 1, f[0]), 1), B("u1", 1, f[0]), 1)
    out = ops.head_conv(u1, pk["outconv.w"], pk["outconv.bias"])
    srcs = dict(zip([hname for hname, _ in _V1_HEADS], (bottom, b2, u3, u2)))
    deeps: List[torch.Tensor] = []
    if want_deep and net.deep_supervision:
        for hname, s in _V1_HEADS:
            deeps.append(ops.upsample_f32(ops.head_conv(srcs[hname], pk[hname + ".w"], pk[hname + ".bias"]), s))
    tape["__meta__"] = dict(shape=(n, d, h, w), ws=ws, cats=(cat1, cat2, cat3, cat4), srcs=srcs, u1=u1,
                            want_deep=bool(deeps))
    return out, deeps, tape


def _backward_v1(net, tape, dout, ddeeps, gs: GradStore):
    pk, G = net._packed, gs.views
    meta = tape["__meta__"]
    n, d, h, w = meta["shape"]
    ws, f = meta["ws"], net.features
    B = lambda name, s, c: net._buf(ws, "g." + name, (n, d // s, h // s, w // s, c))  # noqa: E731
    cat1, cat2, cat3, cat4 = meta["cats"]
    srcs = meta["srcs"]
    nbw = ops.norm_bwd_workspace(n, 2 * f[3], cat1.device)
    dl = dict(zip([hname for hname, _ in _V1_HEADS], ddeeps)) if meta["want_deep"] else {}

    def cbr_bwd(name, dy, dx_out):
        """dy is overwritten with the gradient of the pre-norm conv output."""
        t = tape[name]
        ops.norm_bwd(dy, t["z"], dy, t["st"], pk[name + ".g"], pk[name + ".b"], G[name + ".bn.weight"],
                     G[name + ".bn.bias"], GN, workspace=nbw)
        if dx_out is not None:
            ops.conv3d(dy, pk[name + ".T"], out=dx_out, dil=t["dil"])
        gs.on_side(lambda: ops.conv3d_wgrad(t["x"], dy, G[name + ".conv.weight"], dil=t["dil"]))  # after the dgrad: see V2
        gs.ready(name + ".conv.weight")

    def head_bwd(pname, x, dlog, dx, accumulate):
        dws, db = ops.head_conv_bwd(x, pk[pname + ".w"], dlog, dx, accumulate=accumulate)
        G[pname + ".weight"].view(dws.shape[1], dws.shape[2]).add_(dws.sum(0))
        G[pname + ".bias"].add_(db)
        gs.ready(pname + ".bias")

    def deep_bwd(hname, s, dx):
        g = dl.get(hname)
        if g is not None:
            head_bwd(hname, srcs[hname], ops.upsample_f32_bwd(g.float(), s), dx, accumulate=True)

    if dout is None:
        dout = torch.zeros((n, net.num_classes, d, h, w), dtype=torch.float32, device=cat1.device)
    g_u1 = B("u1", 1, f[0])
    head_bwd("outconv", meta["u1"], dout.float(), g_u1, accumulate=False)
    g_d1, g_cat1 = B("d1", 1, f[0]), B("cat1", 1, 2 * f[0])
    cbr_bwd("decoder1.ConvBnRelu2", g_u1, g_d1)
    cbr_bwd("decoder1.ConvBnRelu1", g_d1, g_cat1)
    g_u2 = B("u2", 2, f[0])
    ops.upsample2x_bwd(g_cat1[..., f[0]:], g_u2)
    deep_bwd("deep2.0", 2, g_u2)
    g_d2, g_cat2 = B("d2", 2, f[1]), B("cat2", 2, 2 * f[1])
    cbr_bwd("decoder2.ConvBnRelu2", g_u2, g_d2)
    cbr_bwd("decoder2.ConvBnRelu1", g_d2, g_cat2)
    g_u3 = B("u3", 4, f[1])
    ops.upsample2x_bwd(g_cat2[..., f[1]:], g_u3)
    deep_bwd("deep3.0", 4, g_u3)
    g_d3, g_cat3 = B("d3", 4, f[2]), B("cat3", 4, 2 * f[2])
    cbr_bwd("decoder3.ConvBnRelu2", g_u3, g_d3)
    cbr_bwd("decoder3.ConvBnRelu1", g_d3, g_cat3)
    g_b2 = B("b2", 8, f[2])
    ops.upsample2x_bwd(g_cat3[..., f[2]:], g_b2)
    deep_bwd("deep_bottom2.0", 8, g_b2)
    g_cat4 = B("cat4", 8, 2 * f[3])
    cbr_bwd("bottom_2", g_b2, g_cat4)
    deep_bwd("deep_bottom.0", 8, g_cat4[..., f[3]:])
    g_bt, g_tmp = B("bt", 8, f[3]), B("tmp4", 8, f[3])
    cbr_bwd("bottom.ConvBnRelu2", g_cat4[..., f[3]:], g_bt)
    cbr_bwd("bottom.ConvBnRelu1", g_bt, g_tmp)
    ops.add_inplace(g_cat4[..., :f[3]], g_tmp)
    g_e4, g_p3 = B("e4", 8, f[3]), B("p3", 8, f[2])
    cbr_bwd("encoder4.ConvBnRelu2", g_cat4[..., :f[3]], g_e4)
    cbr_bwd("encoder4.ConvBnRelu1", g_e4, g_p3)
    ops.pool_bwd(cat3[..., :f[2]], g_p3, g_cat3[..., :f[2]], 1, add=g_cat3[..., :f[2]])
    g_e3, g_p2 = B("e3", 4, f[2]), B("p2", 4, f[1])
    cbr_bwd("encoder3.ConvBnRelu2", g_cat3[..., :f[2]], g_e3)
    cbr_bwd("encoder3.ConvBnRelu1", g_e3, g_p2)
    ops.pool_bwd(cat2[..., :f[1]], g_p2, g_cat2[..., :f[1]], 1, add=g_cat2[..., :f[1]])
    g_e2, g_p1 = B("e2", 2, f[1]), B("p1", 2, f[0])
    cbr_bwd("encoder2.ConvBnRelu2", g_cat2[..., :f[1]], g_e2)
    cbr_bwd("encoder2.ConvBnRelu1", g_e2, g_p1)
    ops.pool_bwd(cat1[..., :f[0]], g_p1, g_cat1[..., :f[0]], 1, add=g_cat1[..., :f[0]])
    g_e1 = B("e1", 1, f[0])
    cbr_bwd("encoder1.ConvBnRelu2", g_cat1[..., :f[0]], g_e1)
    cbr_bwd("encoder1.ConvBnRelu1", g_e1, None)
