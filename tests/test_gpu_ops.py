"""GPU parity of each kernel family (through the C-ABI) against the oracle / plain torch fp32 on identical inputs.
Tolerances: bf16 storage => 2^-8 relative per rounding; integer/byte outputs are bit-exact."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

DEV = "cuda"


@pytest.fixture(autouse=True)
def _no_tf32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


def _cl(x):  # NCDHW -> channels-last bf16
    return x.permute(0, 2, 3, 4, 1).contiguous().to(torch.bfloat16)


def _nc(x):  # channels-last -> NCDHW fp32
    return x.float().permute(0, 4, 1, 2, 3).contiguous()


@pytest.mark.parametrize("cin,cout,k,dil,shape", [
    (48, 48, 3, 1, (1, 16, 16, 16)), (8, 48, 3, 1, (2, 8, 16, 8)), (96, 96, 3, 1, (1, 8, 8, 16)),
    (384, 384, 3, 2, (1, 8, 8, 8)), (384, 96, 3, 6, (1, 16, 16, 16)), (768, 192, 3, 1, (1, 4, 8, 8)),
    (48, 24, 1, 1, (1, 16, 8, 8)), (16, 16, 3, 1, (1, 5, 7, 9)), (64, 64, 3, 4, (1, 2, 2, 2)),
    # two march items per CTA (256 tiles on 148 SMs), 20 planes each: the accumulator ring starts the second item at
    # slot 4 and passes its aliased overflow slots several times
    (48, 48, 3, 1, (4, 20, 128, 64)), (8, 48, 3, 1, (4, 21, 128, 64)), (32, 64, 3, 1, (4, 19, 64, 128))])
def test_conv3d_matches_torch(cin, cout, k, dil, shape):
    from brats21_b200 import ops
    g = torch.Generator(device=DEV).manual_seed(cin * 1000 + cout + dil)
    n, d, h, w = shape
    x = torch.randn((n, cin, d, h, w), device=DEV, generator=g)
    wt = torch.randn((cout, cin, k, k, k), device=DEV, generator=g) / (cin * k ** 3) ** 0.5
    b = torch.randn((cout,), device=DEV, generator=g)
    xb = _cl(x)
    st = ops.new_stats(n, DEV)
    y = ops.conv3d(xb, ops.PackedConv(wt, b), stats=st, dil=dil)
    ref = F.conv3d(xb.float().permute(0, 4, 1, 2, 3), wt.to(torch.bfloat16).float(), b,
                   padding=dil if k == 3 else 0, dilation=dil)
    assert (_nc(y) - ref).abs().max().item() <= 2 ** -7 * ref.abs().max().item()  # one bf16 rounding of the output
    r = ref.reshape(n, 8, cout // 8, -1).double()
    s = st.sum(0)
    assert torch.allclose(s[..., 0], r.sum(dim=(2, 3)), rtol=1e-4, atol=1e-2)
    assert torch.allclose(s[..., 1], (r * r).sum(dim=(2, 3)), rtol=1e-4, atol=1e-2)


@pytest.mark.parametrize("cin,cout,shape", [
    (48, 24, (4, 16, 16, 24)), (96, 48, (2, 5, 7, 9)), (96, 24, (1, 8, 8, 8)), (192, 96, (2, 8, 8, 8)),
    (192, 48, (1, 3, 5, 7)), (384, 96, (1, 8, 8, 8)), (16, 8, (3, 4, 6, 10)), (128, 32, (1, 9, 9, 9)),
    (64, 16, (2, 40, 40, 40))])
def test_conv1x1_persistent_matches_torch_and_tap_kernel(cin, cout, shape):
    """conv_point.cu (persistent 1x1) against torch fp32 and against the tap kernel, writing into a channel slice of
    a wider buffer (the concat-free layout) with ragged tile counts and several samples per launch."""
    from brats21_b200 import _lib, ops
    assert _lib.load().b21_conv_point_supported(cin, cout)
    g = torch.Generator(device=DEV).manual_seed(cin * 7 + cout)
    n, d, h, w = shape
    x = torch.randn((n, cin, d, h, w), device=DEV, generator=g)
    wt = torch.randn((cout, cin, 1, 1, 1), device=DEV, generator=g) / cin ** 0.5
    b = torch.randn((cout,), device=DEV, generator=g)
    xwide = torch.zeros((n, d, h, w, cin + 8), device=DEV, dtype=torch.bfloat16)
    xwide[..., 8:] = _cl(x)
    xb = xwide[..., 8:]
    pw = ops.PackedConv(wt, b)
    assert pw.point_ok
    ywide = torch.full((n, d, h, w, cout + 16), 7.0, device=DEV, dtype=torch.bfloat16)
    st = ops.new_stats(n, DEV)
    y = ops.conv3d(xb, pw, out=ywide[..., 8:8 + cout], stats=st)
    ref = F.conv3d(xb.float().permute(0, 4, 1, 2, 3), wt.to(torch.bfloat16).float(), b)
    assert (_nc(y) - ref).abs().max().item() <= 2 ** -7 * ref.abs().max().item()
    assert (ywide[..., :8] == 7).all() and (ywide[..., 8 + cout:] == 7).all()  # neighbours of the slice untouched
    r = ref.reshape(n, 8, cout // 8, -1).double()
    s = st.sum(0)
    assert torch.allclose(s[..., 0], r.sum(dim=(2, 3)), rtol=1e-4, atol=1e-2)
    assert torch.allclose(s[..., 1], (r * r).sum(dim=(2, 3)), rtol=1e-4, atol=1e-2)
    ops.use_point = False
    try:
        st2 = ops.new_stats(n, DEV)
        y2 = ops.conv3d(xb, pw, stats=st2)
    finally:
        ops.use_point = True
    assert torch.equal(y2, y.contiguous())  # same bf16 inputs, fp32 accumulation over K <= 384: identical roundings
    y3 = ops.conv3d(xb, ops.PackedConv(wt, None))  # no bias, no stats
    assert (_nc(y3) - (ref - b.reshape(1, -1, 1, 1, 1))).abs().max().item() <= 2 ** -7 * ref.abs().max().item()


@pytest.mark.parametrize("cin,cout,shape", [
    (96, 96, (1, 8, 16, 8)), (96, 96, (2, 7, 20, 12)), (96, 96, (1, 32, 32, 32)), (48, 96, (1, 9, 16, 16)),
    (96, 48, (1, 10, 8, 24)), (64, 64, (2, 5, 9, 11)), (96, 192, (1, 4, 16, 8)), (96, 96, (1, 3, 8, 8)),
    (96, 48, (1, 13, 17, 9)),
    # channel-chunked mode (Cin >= 128: 64-channel chunks, accumulators persist across the chunk passes)
    (192, 192, (1, 8, 16, 16)), (192, 192, (2, 7, 9, 12)), (384, 384, (1, 4, 8, 8)), (192, 96, (1, 5, 16, 8)),
    (384, 192, (1, 6, 8, 16)), (128, 64, (1, 10, 8, 8)), (768, 192, (1, 3, 8, 8))])
def test_conv3d_slide_matches_torch(cin, cout, shape):
    """conv_slide.cu (smem-resident halo planes, streamed weight taps, 3 output planes per tap) against torch fp32,
    reading/writing channel slices of wider buffers, with ragged tiles, several d-segments and samples."""
    from brats21_b200 import ops
    g = torch.Generator(device=DEV).manual_seed(cin * 31 + cout)
    n, d, h, w = shape
    x = torch.randn((n, cin, d, h, w), device=DEV, generator=g)
    wt = torch.randn((cout, cin, 3, 3, 3), device=DEV, generator=g) / (cin * 27) ** 0.5
    b = torch.randn((cout,), device=DEV, generator=g)
    xwide = torch.zeros((n, d, h, w, cin + 16), device=DEV, dtype=torch.bfloat16)
    xwide[..., 8:8 + cin] = _cl(x)
    xwide[..., :8] = 3.0
    xwide[..., 8 + cin:] = -5.0  # neighbours of the input slice must not leak in
    xb = xwide[..., 8:8 + cin]
    pw = ops.PackedConv(wt, b)
    assert pw.w_slide is not None and pw.w_march is None
    ywide = torch.full((n, d, h, w, cout + 16), 7.0, device=DEV, dtype=torch.bfloat16)
    st = ops.new_stats(n, DEV)
    y = ops.conv3d(xb, pw, out=ywide[..., 8:8 + cout], stats=st)
    ref = F.conv3d(xb.float().permute(0, 4, 1, 2, 3), wt.to(torch.bfloat16).float(), b, padding=1)
    assert (_nc(y) - ref).abs().max().item() <= 2 ** -7 * ref.abs().max().item()
    assert (ywide[..., :8] == 7).all() and (ywide[..., 8 + cout:] == 7).all()
    r = ref.reshape(n, 8, cout // 8, -1).double()
    s = st.sum(0)
    assert torch.allclose(s[..., 0], r.sum(dim=(2, 3)), rtol=1e-4, atol=1e-2)
    assert torch.allclose(s[..., 1], (r * r).sum(dim=(2, 3)), rtol=1e-4, atol=1e-2)
    # data-gradient packing (transposed + mirrored) == conv_transpose3d
    pt = ops.PackedConv(wt, None, transpose_flip=True)
    if pt.w_slide is not None:
        dy = _cl(torch.randn((n, cout, d, h, w), device=DEV, generator=g))
        dx = ops.conv3d(dy, pt)
        ref_dx = F.conv_transpose3d(dy.float().permute(0, 4, 1, 2, 3), wt.to(torch.bfloat16).float(), padding=1)
        assert (_nc(dx) - ref_dx).abs().max().item() <= 2 ** -7 * ref_dx.abs().max().item()


@pytest.mark.parametrize("cin,cout,k,shape", [
    (48, 48, 3, (2, 6, 16, 16)), (96, 96, 3, (3, 7, 16, 24)), (16, 16, 3, (1, 2, 8, 8)), (64, 64, 3, (2, 5, 9, 11)),
    (48, 24, 1, (3, 5, 6, 7)), (96, 48, 1, (2, 4, 4, 4)), (32, 32, 3, (4, 9, 10, 8))])
def test_conv_fold_matches_explicit_affine(cin, cout, k, shape):
    """Folded EvoNorm (csrc/fold.cu): conv with per-sample weights W*A and the border-class bias table on a stored
    tensor S must equal the explicit formulation swish(conv(zero-padded A*S + B) + bias); statistics are those of the
    pre-activation; channel sums are those of the stored outputs."""
    from brats21_b200 import ops
    g = torch.Generator(device=DEV).manual_seed(cin + 13 * cout + k)
    n, d, h, w = shape
    s_in = _cl(torch.randn((n, cin, d, h, w), device=DEV, generator=g))
    a = torch.randn((n, cin), device=DEV, generator=g) * 0.5 + 1.0
    a[:, ::5] *= -1.0
    b = torch.randn((n, cin), device=DEV, generator=g) * 0.7
    wt = torch.randn((cout, cin, k, k, k), device=DEV, generator=g) / (cin * k ** 3) ** 0.5
    bias = torch.randn((cout,), device=DEV, generator=g)
    pw = ops.PackedConv(wt, bias)
    assert ops.fold_supported(pw)
    ab = torch.stack([a, b])  # [2, n, cin]
    st = ops.new_stats(n, DEV)
    csum = torch.zeros((n, cout), device=DEV) if k == 3 else None
    out = torch.empty((n, d, h, w, cout), device=DEV, dtype=torch.bfloat16)
    ops.conv3d_fold(s_in, pw, out, st, ab=(ab[0], ab[1]), act=True, chan_sum=csum)
    # explicit reference in fp32 with the SAME bf16-rounded folded weights the kernel uses
    y_in = _nc(s_in) * 1.0
    refs = []
    for i in range(n):
        wq = (wt * a[i].reshape(1, cin, 1, 1, 1)).to(torch.bfloat16).float()
        z = F.conv3d(y_in[i:i + 1], wq, None, padding=k // 2)
        ones = torch.ones((1, cin, d, h, w), device=DEV) * b[i].reshape(1, cin, 1, 1, 1)
        z = z + F.conv3d(ones, wt, bias, padding=k // 2)  # response to the constant image B incl. the zero-padded border
        refs.append(z)
    z = torch.cat(refs)
    ref = z * torch.sigmoid(z)
    assert (_nc(out) - ref).abs().max().item() <= 2 ** -7 * ref.abs().max().item() + 1e-3
    r = z.reshape(n, 8, cout // 8, -1).double()
    sm = st.sum(0)
    assert torch.allclose(sm[..., 0], r.sum(dim=(2, 3)), rtol=1e-3, atol=5e-2)
    assert torch.allclose(sm[..., 1], (r * r).sum(dim=(2, 3)), rtol=1e-3, atol=5e-2)
    if csum is not None:
        assert torch.allclose(csum, _nc(out).sum(dim=(2, 3, 4)), rtol=1e-3, atol=5e-2)
    # shared weights (no input affine), no activation == the plain conv
    out2 = torch.empty_like(out)
    ops.conv3d_fold(s_in, pw, out2, st, ab=None, act=False)
    assert torch.equal(out2, ops.conv3d(s_in, pw))


@pytest.mark.parametrize("cin,cout,shape", [
    (4, 48, (2, 8, 16, 8)), (4, 48, (1, 5, 7, 9)), (4, 16, (3, 1, 8, 8)), (3, 32, (1, 2, 20, 13)),
    (1, 64, (2, 11, 17, 24)),
    # two items per CTA (256 tiles on 148 SMs), odd plane count: block-pair stages and accumulator buffers wrap many times
    (4, 48, (4, 21, 128, 64)), (4, 48, (1, 128, 128, 128))])
def test_conv_input_matches_torch_and_march(cin, cout, shape):
    """First-conv kernel (csrc/conv_input.cu: im2col plane blocks built with cp.async, pairs of planes per hand-shake):
    same result as F.conv3d on the bf16-rounded operands, statistics of the pre-activation, swish variants, and the
    same values as the plane-marching kernel it replaces."""
    from brats21_b200 import ops
    g = torch.Generator(device=DEV).manual_seed(cin * 100 + cout + shape[1])
    n, d, h, w = shape
    x = torch.randn((n, cin, d, h, w), device=DEV, generator=g)
    wt = torch.randn((cout, cin, 3, 3, 3), device=DEV, generator=g) / (cin * 27) ** 0.5
    b = torch.randn((cout,), device=DEV, generator=g)
    xb = torch.zeros((n, d, h, w, 8), device=DEV, dtype=torch.bfloat16)
    xb[..., :cin] = _cl(x)
    pw = ops.PackedConv(wt, b)
    assert pw.w_input is not None and pw.cin == 8
    ref = F.conv3d(xb[..., :cin].float().permute(0, 4, 1, 2, 3), wt.to(torch.bfloat16).float(), b, padding=1)
    saved = ops.use_input
    try:
        ops.use_input = True
        st = ops.new_stats(n, DEV)
        y = ops.conv3d(xb, pw, stats=st)
        assert (_nc(y) - ref).abs().max().item() <= 2 ** -7 * ref.abs().max().item()
        r = ref.reshape(n, 8, cout // 8, -1).double()
        sm = st.sum(0)
        assert torch.allclose(sm[..., 0], r.sum(dim=(2, 3)), rtol=1e-4, atol=1e-2)
        assert torch.allclose(sm[..., 1], (r * r).sum(dim=(2, 3)), rtol=1e-4, atol=1e-2)
        sw = ref * torch.sigmoid(ref)
        for fast in (False, True):
            ops.fast_input_swish = fast
            out = torch.empty_like(y)
            st2 = ops.new_stats(n, DEV)
            ops.conv3d_fold(xb, pw, out, st2, ab=None, act=True)
            assert (_nc(out) - sw).abs().max().item() <= 2 ** -7 * sw.abs().max().item() + 2e-3
            assert torch.allclose(st2.sum(0), sm, rtol=1e-6, atol=1e-6)
        ops.use_input = False
        if h >= 8 and w >= 8:
            y_march = ops.conv3d(xb, pw)
            assert (y.float() - y_march.float()).abs().max().item() <= 2 ** -7 * ref.abs().max().item()
    finally:
        ops.use_input = saved
        ops.fast_input_swish = True


def test_evo_se_affine_and_affine_pool():
    from brats21_b200 import ops
    from oracle import nets
    g = torch.Generator(device=DEV).manual_seed(11)
    n, c, d, h, w = 3, 32, 4, 6, 8
    x = torch.randn((n, c, d, h, w), device=DEV, generator=g) * 1.3 + 0.2
    gamma = 1 + 0.3 * torch.randn(c, device=DEV, generator=g)
    gamma[3] = -0.8
    beta = 0.2 * torch.randn(c, device=DEV, generator=g)
    w1, b1 = torch.randn((c // 2, c), device=DEV, generator=g) * 0.3, torch.randn(c // 2, device=DEV, generator=g) * 0.1
    w2, b2 = torch.randn((c, c // 2), device=DEV, generator=g) * 0.3, torch.randn(c, device=DEV, generator=g) * 0.1
    st = torch.zeros((32, n, 8, 2), dtype=torch.float64, device=DEV)
    r = x.reshape(n, 8, -1).double()
    st[1, :, :, 0] = r.sum(-1)
    st[5, :, :, 1] = (r * r).sum(-1)
    sw = _cl(x * torch.sigmoid(x))           # what a folded conv epilogue stores
    csum = _nc(sw).sum(dim=(2, 3, 4))
    ab = torch.zeros((2, n, c + 8), device=DEV)
    ops.evo_se_affine(st, gamma, beta, ab[0][:, 8:], ab[1][:, 8:], d * h * w, chan_sum=csum, se=(w1, b1, w2, b2))
    y = nets.evonorm_s0(x, gamma, beta)
    ref = nets.residual_se(y, w1, b1, w2, b2)
    got = _nc(sw) * ab[0][:, 8:].reshape(n, c, 1, 1, 1) + ab[1][:, 8:].reshape(n, c, 1, 1, 1)
    assert (got - ref).abs().max().item() <= 2 ** -7 * ref.abs().max().item() + 1e-3
    assert (ab[:, :, :8] == 0).all()
    # without SE
    ab2 = torch.zeros((2, n, c), device=DEV)
    ops.evo_se_affine(st, gamma, beta, ab2[0], ab2[1], d * h * w)
    got2 = _nc(sw) * ab2[0].reshape(n, c, 1, 1, 1) + ab2[1].reshape(n, c, 1, 1, 1)
    assert (got2 - y).abs().max().item() <= 2 ** -7 * y.abs().max().item() + 1e-3
    # MaxAvgPool of the affine tensor (negative A -> min)
    pooled = torch.zeros((n, d // 2, h // 2, w // 2, 2 * c), device=DEV, dtype=torch.bfloat16)
    ops.affine_pool(sw, ab[0][:, 8:], ab[1][:, 8:], pooled, mode=2)
    ref_pool = nets.max_avg_pool(got)
    assert (_nc(pooled) - ref_pool).abs().max().item() <= 2 ** -7 * ref_pool.abs().max().item() + 1e-3
    # head with scale + offset == head on the explicit tensor
    wh, bh = torch.randn((3, c), device=DEV, generator=g) * 0.2, torch.randn(3, device=DEV, generator=g)
    lo = ops.head_conv(sw, wh, bh, scale=ab[0][:, 8:], offset=ab[1][:, 8:])
    ref_lo = F.conv3d(got, wh.reshape(3, c, 1, 1, 1), bh)
    assert torch.allclose(lo, ref_lo, rtol=1e-4, atol=1e-4)


def test_conv3d_argument_errors():
    from brats21_b200 import ops
    x = torch.zeros((1, 4, 4, 4, 12), device=DEV, dtype=torch.bfloat16)
    pw = ops.PackedConv(torch.zeros((8, 16, 1, 1, 1), device=DEV), None)
    with pytest.raises((RuntimeError, AssertionError)):
        ops.conv3d(x, pw)


@pytest.mark.parametrize("mode", ["gn", "evo"])
@pytest.mark.parametrize("c,shape", [(48, (2, 8, 8, 8)), (16, (1, 4, 6, 10)), (384, (1, 4, 4, 4))])
def test_norm_apply_matches_oracle(mode, c, shape):
    from brats21_b200 import ops
    from oracle import nets
    g = torch.Generator(device=DEV).manual_seed(c)
    n, d, h, w = shape
    x = torch.randn((n, c, d, h, w), device=DEV, generator=g) * 1.7 + 0.3
    gamma = 1 + 0.2 * torch.randn(c, device=DEV, generator=g)
    beta = 0.2 * torch.randn(c, device=DEV, generator=g)
    xb = _cl(x)
    xq = _nc(xb)
    # statistics exactly as the conv epilogue would deliver them (fp64 sums of the fp32 values)
    st = torch.zeros((32, n, 8, 2), dtype=torch.float64, device=DEV)
    r = xq.reshape(n, 8, -1).double()
    st[3, :, :, 0] = r.sum(-1)
    st[7, :, :, 1] = (r * r).sum(-1)
    csum = torch.zeros((n, c), device=DEV)
    out = torch.empty_like(xb)
    ops.norm_apply(xb, st, gamma, beta, ops.GN_RELU if mode == "gn" else ops.EVO_S0, out=out, chan_sum=csum)
    ref = nets.group_norm_relu(xq, gamma, beta) if mode == "gn" else nets.evonorm_s0(xq, gamma, beta)
    assert (_nc(out) - ref).abs().max().item() <= 2 ** -7 * ref.abs().max().item() + 1e-3
    assert torch.allclose(csum, _nc(out).sum(dim=(2, 3, 4)), rtol=1e-3, atol=1e-2)
    # in place
    ops.norm_apply(xb, st, gamma, beta, ops.GN_RELU if mode == "gn" else ops.EVO_S0)
    assert torch.equal(xb, out)


def test_se_gate_scale_pool_upsample_head():
    from brats21_b200 import ops
    from oracle import nets
    g = torch.Generator(device=DEV).manual_seed(5)
    n, c, d, h, w = 2, 32, 8, 12, 16
    x = torch.randn((n, c, d, h, w), device=DEV, generator=g)
    xb = _cl(x)
    xq = _nc(xb)
    w1, b1 = torch.randn((c // 2, c), device=DEV, generator=g) * 0.3, torch.randn(c // 2, device=DEV, generator=g) * 0.1
    w2, b2 = torch.randn((c, c // 2), device=DEV, generator=g) * 0.3, torch.randn(c, device=DEV, generator=g) * 0.1
    csum = xq.sum(dim=(2, 3, 4))
    scale = ops.se_gate(csum, w1, b1, w2, b2, d * h * w)
    ref_scale = 1 + torch.sigmoid(F.linear(torch.relu(F.linear(xq.mean(dim=(2, 3, 4)), w1, b1)), w2, b2))
    assert torch.allclose(scale, ref_scale, rtol=1e-5, atol=1e-5)
    se_ref = nets.residual_se(xq, w1, b1, w2, b2)
    full = torch.empty_like(xb)
    pooled = torch.zeros((n, d // 2, h // 2, w // 2, 2 * c + 8), device=DEV, dtype=torch.bfloat16)
    ops.scale_pool(xb, scale, full=full, pooled=pooled[..., :2 * c], mode=2)
    assert (_nc(full) - se_ref).abs().max().item() <= 2 ** -7 * se_ref.abs().max().item()
    ref_pool = nets.max_avg_pool(_nc(full))
    assert (_nc(pooled[..., :2 * c]) - ref_pool).abs().max().item() <= 2 ** -7 * ref_pool.abs().max().item()
    assert (pooled[..., 2 * c:] == 0).all()
    mp = torch.empty((n, d // 2, h // 2, w // 2, c), device=DEV, dtype=torch.bfloat16)
    ops.scale_pool(xb, None, pooled=mp, mode=1)
    assert torch.equal(_nc(mp), F.max_pool3d(xq, 2))
    up = torch.zeros((n, 2 * d, 2 * h, 2 * w, 2 * c), device=DEV, dtype=torch.bfloat16)
    ops.upsample2x(xb, up[..., c:])
    ref_up = nets.up_trilinear(xq, 2)
    assert (_nc(up[..., c:]) - ref_up).abs().max().item() <= 2 ** -7 * ref_up.abs().max().item()
    hw, hb = torch.randn((3, c), device=DEV, generator=g) * 0.2, torch.randn(3, device=DEV, generator=g)
    logits = ops.head_conv(xb, hw, hb, scale=scale)
    ref_logits = F.conv3d(xq * scale.reshape(n, c, 1, 1, 1), hw.reshape(3, c, 1, 1, 1), hb)
    assert torch.allclose(logits, ref_logits, rtol=1e-4, atol=1e-4)
    for s in (2, 4, 8):
        small = logits[:, :, :4, :6, :8].contiguous()
        assert torch.allclose(ops.upsample_f32(small, s), nets.up_trilinear(small, s), rtol=1e-5, atol=1e-5)


def test_pack_blend_tta_labels_bit_exact():
    from brats21_b200 import ops, tta
    from oracle import inference as oinf
    g = torch.Generator(device=DEV).manual_seed(9)
    vol = torch.randn((2, 4, 12, 10, 14), device=DEV, generator=g)
    for tr in list(tta.get_tta_transforms()) + list(tta.get_flip8_transforms()):
        perm, flip = tr.variant
        aug = tr.augment_image(vol).contiguous()
        ad = aug.shape[2:]
        out = torch.empty((2, 8, 8, 8, 8), device=DEV, dtype=torch.bfloat16)
        origins = [(1, 0, 2), (-3, 2, ad[2] - 6)]  # second window hangs over both ends -> zero padding
        ops.pack_windows(vol, out, origins, perm=perm, flip=flip, vol_index=[0, 1])
        for b, o in enumerate(origins):
            ref = torch.zeros((4, 8, 8, 8), device=DEV)
            lo = [max(0, -c) for c in o]
            hi = [min(8, ad[j] - o[j]) for j in range(3)]
            ref[:, lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]] = aug[b, :, o[0] + lo[0]:o[0] + hi[0], o[1] + lo[1]:o[1] + hi[1],
                                                                o[2] + lo[2]:o[2] + hi[2]]
            got = out[b].float().permute(3, 0, 1, 2)
            assert torch.equal(got[:4], ref.to(torch.bfloat16).float())
            assert (got[4:] == 0).all()
        # de-augmentation: prob_sum += sigmoid(acc / cnt) scattered back
        acc = torch.randn((3,) + tuple(ad), device=DEV, generator=g)
        cnt = torch.rand((1,) + tuple(ad), device=DEV, generator=g) + 0.5
        ps = torch.full((3, 12, 10, 14), 0.25, device=DEV)
        ops.tta_accumulate(acc, cnt, ps, perm, flip)
        ref = 0.25 + tr.deaugment_mask(torch.sigmoid(acc / cnt)[None])[0]
        assert torch.allclose(ps, ref, rtol=1e-6, atol=1e-6)
    # blending: fp32, same operation order as the reference loop => bit-exact against torch on the GPU
    prof = [torch.rand(8, device=DEV, generator=g) + 0.1 for _ in range(3)]
    logits = torch.randn((3, 3, 8, 8, 8), device=DEV, generator=g)
    acc = torch.zeros((3, 12, 10, 14), device=DEV)
    origins = [(0, 0, 0), (4, 2, 6), (2, 1, 3)]
    ops.blend_accumulate(logits, acc, prof, origins)
    ref = torch.zeros_like(acc)
    wmap = prof[0].reshape(-1, 1, 1) * prof[1].reshape(1, -1, 1) * prof[2].reshape(1, 1, -1)
    for j, o in enumerate(origins):
        ref[:, o[0]:o[0] + 8, o[1]:o[1] + 8, o[2]:o[2] + 8] += wmap * logits[j]
    assert torch.equal(acc, ref)
    with pytest.raises(RuntimeError):
        ops.blend_accumulate(logits, acc, prof, [(0, 0, 0), (8, 2, 6), (2, 1, 3)])
    # labels: bit-exact against the oracle's post-processing
    prob_sum = torch.rand((3, 12, 10, 14), device=DEV, generator=g) * 8
    img = vol[0].clone()
    img[:, :3] = 0
    onehot, label = ops.labels_finalize(prob_sum, 8, 0.5, image=img)
    hard = ((prob_sum / 8) >= 0.5).float()[None].cpu()
    hard = oinf.remove_background_voxels(img[None].cpu(), hard)
    assert torch.equal(onehot.cpu().float()[None], hard)
    assert torch.equal(label.cpu()[None, None], oinf.brats_label_map(hard))


def test_blend_and_tta_vectorised_paths():
    """float4 paths (x extents / origins multiples of 4; flip-only variants): blending bit-exact in the reference's
    window order, de-augmentation against torch flips, and MONAI's importance-map floor for sigma_scale < 1/8."""
    from brats21_b200 import inferers, ops
    from oracle import inference as oinf
    g = torch.Generator(device=DEV).manual_seed(17)
    prof = [torch.rand(8, device=DEV, generator=g) + 0.1 for _ in range(3)]
    logits = torch.randn((4, 3, 8, 8, 8), device=DEV, generator=g)
    acc = torch.randn((3, 16, 12, 16), device=DEV, generator=g)
    ref = acc.clone()
    origins = [(0, 0, 0), (0, 4, 4), (8, 4, 8), (6, 3, 4)]  # overlapping, all x origins multiples of 4
    ops.blend_accumulate(logits, acc, prof, origins)
    wmap = prof[0].reshape(-1, 1, 1) * prof[1].reshape(1, -1, 1) * prof[2].reshape(1, 1, -1)
    for j, o in enumerate(origins):
        ref[:, o[0]:o[0] + 8, o[1]:o[1] + 8, o[2]:o[2] + 8] += wmap * logits[j]
    assert torch.equal(acc, ref)
    # every flip subset, padded augmented frame, accumulate and overwrite
    a = torch.randn((3, 14, 12, 24), device=DEV, generator=g)
    cnt = torch.rand((1, 14, 12, 24), device=DEV, generator=g) + 0.5
    for bits in range(8):
        flip = (bits & 1, (bits >> 1) & 1, (bits >> 2) & 1)
        ps = torch.full((3, 12, 10, 16), 0.5, device=DEV)
        ops.tta_accumulate(a, cnt, ps, (0, 1, 2), flip, pad_before=(1, 2, 4))
        core = torch.sigmoid(a / cnt)[:, 1:13, 2:12, 4:20]
        dims = [i + 1 for i in range(3) if flip[i]]
        want = 0.5 + (core.flip(dims) if dims else core)
        assert torch.allclose(ps, want, rtol=1e-6, atol=1e-6), flip
        ops.tta_accumulate(a, cnt, ps, (0, 1, 2), flip, pad_before=(1, 2, 4), apply_sigmoid=False, overwrite=True)
        core = (a / cnt)[:, 1:13, 2:12, 4:20]
        assert torch.equal(ps, core.flip(dims) if dims else core), flip
    # sigma_scale 0.05 on a 32-voxel window: the truncated Gaussian has zeros inside the window; MONAI clamps the 3-D
    # map at its smallest non-zero value, so the count map stays positive and out/count finite
    roi, img = (32, 32, 32), (40, 32, 48)
    plan = inferers.WindowPlan(img, roi, 0.25, "gaussian", 0.05, torch.device(DEV))
    assert plan.wfloor > 0 and plan.count.min().item() > 0
    wm = oinf.importance_map(roi, "gaussian", 0.05).to(DEV)
    want = torch.zeros((1,) + img, device=DEV)
    for o in plan.origins:
        want[:, o[0]:o[0] + 32, o[1]:o[1] + 32, o[2]:o[2] + 32] += wm
    assert torch.allclose(plan.count, want, rtol=1e-6, atol=1e-30)
    x = torch.randn((1, 3) + img, device=DEV, generator=g)
    y = inferers.sliding_window_inference(x, roi, 2, lambda z: z, overlap=0.25, mode="gaussian", sigma_scale=0.05)
    assert torch.isfinite(y).all() and (y - x).abs().max().item() <= 1e-5
