"""GPU parity of the drop-in networks against the oracle (torch fp32 restatement of the reference) and against
the golden vectors produced by the unmodified reference.

Tolerance (stated per BASELINE.json north_star): logits of the bf16 kernels vs the fp32 reference within
relative L2 <= 2.5e-2 and max-abs <= 4e-2 * max|logit|.  For scale: torch's own bf16 autocast vs fp32 on these
random-init networks gives rel-L2 0.8 % (V1) / 1.4 % (V2)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"
REL_L2_TOL, MAX_ABS_TOL = 2.5e-2, 4e-2


@pytest.fixture(autouse=True)
def _no_tf32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


def _build(ver, width, seed):
    import warnings
    from brats21_b200 import networks
    from oracle import synth
    params = {k: v.to(DEV) for k, v in synth.make_params(ver, width, seed).items()}
    cls = networks.EquiUnet if ver == 1 else networks.EquiUnetASSPEvo
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        net = cls(4, 3, [width * 2 ** i for i in range(4)], norm_layer="group", act="relu",
                  deep_supervision=True).to(DEV).eval()
    net.load_state_dict(params, strict=True)
    return net, params


def _check(got, ref):
    got, ref = got.float(), ref.float()
    rel = ((got - ref).norm() / ref.norm()).item()
    mx = (got - ref).abs().max().item() / ref.abs().max().item()
    assert rel <= REL_L2_TOL and mx <= MAX_ABS_TOL, (rel, mx)


@pytest.mark.parametrize("ver,seed", [(1, 123), (2, 93)])
def test_forward_matches_reference_golden(golden_dir, ver, seed):
    from oracle import synth
    g = np.load(os.path.join(golden_dir, f"net_v{ver}_w16.npz"))
    net, _ = _build(ver, 16, seed)
    x = synth.volume(seed=3, shape=(32, 32, 32)).to(DEV)
    out, deeps = net(x)
    assert out.shape == (1, 3, 32, 32, 32) and out.dtype == torch.float32
    _check(out.cpu(), torch.from_numpy(g["out"]))
    assert len(deeps) == (4 if ver == 1 else 2)
    for i, dp in enumerate(deeps):
        assert dp.shape == out.shape
        _check(dp[..., ::2, ::2, ::2].cpu(), torch.from_numpy(g[f"deep{i}_s2"]))


@pytest.mark.parametrize("ver,seed,width,shape,n", [(1, 123, 16, (16, 24, 40), 2), (2, 93, 16, (16, 24, 40), 2),
                                                     (1, 123, 48, (64, 64, 64), 1), (2, 93, 48, (64, 64, 64), 1)])
def test_forward_matches_oracle(ver, seed, width, shape, n):
    from oracle import nets, synth
    net, params = _build(ver, width, seed)
    x = torch.cat([synth.volume(seed=s, shape=shape) for s in range(n)]).to(DEV)
    fwd = nets.equiunet_v1_forward if ver == 1 else nets.equiunet_v2_forward
    with torch.no_grad():
        ref, ref_deeps = fwd(params, x)
    out, deeps = net(x)
    _check(out, ref)
    for a, b in zip(deeps, ref_deeps):
        _check(a, b)
    # Batch independence.  Not bit-exact: the norm statistics are accumulated with atomics, and a 1e-7 change of
    # a statistic is amplified to the bf16 rounding-noise floor within a few layers (each bf16 rounding turns a
    # perturbation d into ~sqrt(d * ulp)), so two runs differ at the same level as bf16 vs fp32.
    out2, _ = net(x[:1].contiguous())
    assert ((out2[0] - net(x)[0][0]).norm() / out2[0].norm()).item() <= 1.5e-2
    # per-region hard-label agreement where the reference margin exceeds the tolerance
    margin = ref.abs() > MAX_ABS_TOL * ref.abs().max()
    assert ((out >= 0) == (ref >= 0))[margin].all()
    inter = ((out >= 0) & (ref >= 0)).sum(dim=(0, 2, 3, 4)).float()
    dice = 2 * inter / ((out >= 0).sum(dim=(0, 2, 3, 4)) + (ref >= 0).sum(dim=(0, 2, 3, 4))).float().clamp_min(1)
    assert (dice >= 0.99).all(), dice


def test_state_dict_round_trip_and_repack():
    net, params = _build(2, 16, 93)
    sd = net.state_dict()
    for k, v in params.items():
        assert torch.equal(sd[k], v)
    from oracle import synth
    x = synth.volume(seed=3, shape=(16, 16, 16)).to(DEV)
    a = net(x)[0].clone()
    net.load_state_dict({k: v.to(DEV) for k, v in synth.make_params(2, 16, 7).items()})
    b = net(x)[0]
    assert not torch.allclose(a, b)  # packed weights follow the parameters
    with pytest.raises(ValueError):
        net(torch.zeros((1, 4, 12, 16, 16), device=DEV))


@pytest.mark.parametrize("width,shape,n", [(16, (16, 32, 48), 2), (48, (64, 64, 64), 1)])
def test_v2_folded_evonorm_path_matches_explicit_path(width, shape, n):
    """The folded-EvoNorm inference path (stored swish tensors, affine folded into the consumers: csrc/fold.cu) and
    the explicit path (norm_apply passes) are two roundings of the same arithmetic: both must sit within the bf16
    tolerance of the fp32 oracle and within 1.5e-2 rel-L2 of each other."""
    from brats21_b200 import ops
    from oracle import nets, synth
    net, params = _build(2, width, 93)
    x = torch.cat([synth.volume(seed=s, shape=shape) for s in range(n)]).to(DEV)
    with torch.no_grad():
        ref, ref_deeps = nets.equiunet_v2_forward(params, x)
    assert ops.use_fold
    net._ensure_packed()
    assert net._fold_ok(*shape)
    out_f, deeps_f = net(x)
    ops.use_fold = False
    try:
        out_e, deeps_e = net(x)
    finally:
        ops.use_fold = True
    _check(out_f, ref)
    _check(out_e, ref)
    for a, b in zip(deeps_f, ref_deeps):
        _check(a, b)
    assert ((out_f - out_e).norm() / out_e.norm()).item() <= 1.5e-2


@pytest.mark.parametrize("width,shape,n", [(16, (16, 32, 48), 2), (48, (32, 48, 40), 1)])
def test_v2_split_concat_matches_in_place_concat(width, shape, n):
    """decoder1's first conv reading [bridge1 | up(upconv1)] as two dense tensors (b21_conv3d_march_fwd_fold2) is
    the same arithmetic on the same values as reading one concat buffer.  The group statistics are accumulated with
    atomics (order varies run to run), so bf16 roundings flip run to run: rel-L2 <= 1e-2 (a wrong channel mapping gives O(1))."""
    from brats21_b200 import ops
    from oracle import synth
    net, _ = _build(2, width, 93)
    x = torch.cat([synth.volume(seed=s, shape=shape) for s in range(n)]).to(DEV)
    assert ops.split_concat
    net._ensure_packed()
    assert net._packed["decoder1.c0"].w_march is not None
    out_s, _ = net(x)
    ops.split_concat = False
    try:
        out_c, _ = net(x)
    finally:
        ops.split_concat = True
    rel = ((out_s - out_c).norm() / out_c.norm()).item()
    assert rel <= 1e-2, rel


@pytest.mark.parametrize("width,shape,n", [(16, (16, 32, 48), 2), (48, (32, 64, 64), 1)])
def test_v2_level3_fold_matches_explicit_level3(width, shape, n):
    """Folding level 3 as well (64- / 192-channel sliding-window convs with per-sample weights, bridge3, the deep3
    head) against the same network with level 3 on the explicit normalisation path and against the fp32 oracle."""
    from brats21_b200 import ops
    from oracle import nets, synth
    net, params = _build(2, width, 93)
    x = torch.cat([synth.volume(seed=s, shape=shape) for s in range(n)]).to(DEV)
    with torch.no_grad():
        ref, ref_deeps = nets.equiunet_v2_forward(params, x)
    assert ops.fold_level3
    out_f, deeps_f = net(x)
    ops.fold_level3 = False
    try:
        out_e, deeps_e = net(x)
    finally:
        ops.fold_level3 = True
    _check(out_f, ref)
    for a, b in zip(deeps_f, ref_deeps):
        _check(a, b)
    assert ((out_f - out_e).norm() / out_e.norm()).item() <= 1.5e-2
    assert ((deeps_f[0] - deeps_e[0]).norm() / deeps_e[0].norm()).item() <= 1.5e-2


@pytest.mark.parametrize("norm,act", [("instance", "leakyrelu"), ("batch", "elu"), ("none", "relu"), ("group", "leakyrelu"),
                                      ("instance", "relu")])
def test_v1_norm_act_factory(golden_dir, norm, act):
    """EquiUnet under the rest of networks/factory.py:179-200 (get_norm_layer instance / batch / none, get_act leakyrelu
    / elu): channel_stats -> norm_coeffs -> affine_act, vs the unmodified-reference golden and the oracle; batch norm in
    eval mode (running statistics) and in training-mode forward (batch statistics, running statistics updated)."""
    import warnings
    from brats21_b200 import networks
    from oracle import nets, synth
    g = np.load(os.path.join(golden_dir, "net_v1_w16_variants.npz"))
    tag = f"{norm}_{act}"
    params = {k: v.to(DEV) for k, v in synth.make_params(1, 16, 123, norm=norm).items()}
    net = networks.EquiUnet(4, 3, [16, 32, 64, 128], norm_layer=norm, act=act, deep_supervision=True).to(DEV).eval()
    assert list(net.state_dict().keys()) == list(g[f"{tag}_keys"])
    net.load_state_dict(params, strict=True)
    x = synth.volume(seed=3, shape=(16, 16, 16)).to(DEV)
    out, deeps = net(x)
    _check(out.cpu(), torch.from_numpy(g[f"{tag}_out"]))
    _check(deeps[0][..., ::2, ::2, ::2].cpu(), torch.from_numpy(g[f"{tag}_deep0_s2"]))
    # a larger, batched, non-cubic input against the oracle
    xb = torch.cat([synth.volume(seed=s, shape=(16, 24, 40)) for s in range(2)]).to(DEV)
    with torch.no_grad():
        ref, _ = nets.equiunet_v1_forward(params, xb, norm=norm, act=act)
    _check(net(xb)[0], ref)
    if norm == "batch":
        net.train()
        with torch.no_grad():
            out_t, _ = net(x)
        _check(out_t.cpu(), torch.from_numpy(g[f"{tag}_train_out"]))
        sd = net.state_dict()
        assert (sd["encoder1.ConvBnRelu1.bn.running_mean"].cpu() -
                torch.from_numpy(g[f"{tag}_train_running_mean"])).abs().max().item() <= 2e-3
        assert (sd["decoder1.ConvBnRelu2.bn.running_var"].cpu() -
                torch.from_numpy(g[f"{tag}_train_running_var"])).abs().max().item() <= 2e-2
        assert int(sd["encoder1.ConvBnRelu1.bn.num_batches_tracked"]) == 1
    if (norm, act) != ("group", "relu"):
        net.train()
        with pytest.raises(NotImplementedError):
            net(x.requires_grad_(True))
    with pytest.raises(ValueError):
        networks.EquiUnet(4, 3, [16, 32, 64, 128], norm_layer=None)
