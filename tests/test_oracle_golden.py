"""CPU: the oracle port (oracle/*.py) against the committed golden vectors that tests/golden/make_golden.py
produced by running the unmodified reference.  Tolerances are fp32 round-off across CPUs (oneDNN kernel choice)."""
import os

import numpy as np
import pytest
import torch

from oracle import inference as oinf
from oracle import nets, synth, train

WIDTH, SHAPE = 16, (32, 32, 32)


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


@pytest.mark.parametrize("ver,seed", [(1, 123), (2, 93)])
def test_network_forward_matches_reference(golden_dir, ver, seed):
    g = _load(golden_dir, f"net_v{ver}_w{WIDTH}.npz")
    x = synth.volume(seed=3, shape=SHAPE)
    p = synth.make_params(ver, WIDTH, seed)
    assert sorted(p.keys()) == list(g["keys"])  # state_dict key contract
    fwd = nets.equiunet_v1_forward if ver == 1 else nets.equiunet_v2_forward
    with torch.no_grad():
        out, deeps = fwd(p, x)
    ref = torch.from_numpy(g["out"])
    scale = ref.abs().max().item()
    assert (out - ref).abs().max().item() <= 2e-4 * scale
    assert len(deeps) == (4 if ver == 1 else 2)
    for i, dp in enumerate(deeps):
        r = torch.from_numpy(g[f"deep{i}_s2"])
        assert (dp[..., ::2, ::2, ::2] - r).abs().max().item() <= 2e-4 * max(r.abs().max().item(), 1.0)


def _ramp_predictor(x):
    d, h, w = x.shape[2:]
    i = torch.arange(d, dtype=x.dtype).reshape(1, d, 1, 1)
    j = torch.arange(h, dtype=x.dtype).reshape(1, 1, h, 1)
    k = torch.arange(w, dtype=x.dtype).reshape(1, 1, 1, w)
    ramp = 1.0 + 0.01 * (i + 2 * j + 3 * k)
    y = torch.stack([x[:, 0] * 0.5 + x[:, 1], x[:, 2] - x[:, 3], x.sum(1) * 0.25], dim=1)
    return y * ramp.unsqueeze(0)


def test_sliding_window_matches_reference(golden_dir):
    g = _load(golden_dir, "sliding_window.npz")
    xs = synth.volume(seed=5, shape=(40, 36, 29))
    for mode in ("constant", "gaussian"):
        for bs in (1, 4):
            y = oinf.sliding_window_inference(xs, (16, 16, 16), bs, _ramp_predictor, overlap=0.25, mode=mode)
            ref = torch.from_numpy(g[f"{mode}_b{bs}"])
            assert y.shape == ref.shape
            assert (y - ref).abs().max().item() <= 1e-5 * ref.abs().max().item()
    y = oinf.sliding_window_inference(xs, (48, 32, 16), 2, _ramp_predictor, overlap=0.5, mode="gaussian")
    ref = torch.from_numpy(g["pad_gaussian"])
    assert y.shape == ref.shape
    assert (y - ref).abs().max().item() <= 1e-5 * ref.abs().max().item()


def test_window_grid_brats_volume():
    # 240x240x155, roi 128, overlap 0.25 -> starts [0,96,112]^2 x [0,27] = 18 windows
    grid = oinf.window_grid((240, 240, 155), (128, 128, 128), 0.25)
    assert len(grid) == 18
    assert sorted({g[0] for g in grid}) == [0, 96, 112]
    assert sorted({g[2] for g in grid}) == [0, 27]
    assert grid[0] == (0, 0, 0) and grid[1] == (0, 0, 27)  # last spatial dim fastest
    assert sorted({g[2] for g in oinf.window_grid((240, 240, 160), (128,) * 3, 0.25)}) == [0, 32]
    with pytest.raises(AssertionError):
        oinf.sliding_window_inference(torch.zeros(1, 1, 8, 8, 8), 4, 1, lambda x: x, overlap=1.0)


def test_tta16_matches_reference(golden_dir):
    g = _load(golden_dir, "tta16.npz")
    variants = oinf.reference_tta()
    assert len(variants) == int(g["n"]) == 16
    xt = torch.arange(2 * 5 * 5 * 5, dtype=torch.float32).reshape(1, 2, 5, 5, 5)
    for i, v in enumerate(variants):
        a = v.augment_image(xt)
        assert np.array_equal(a.contiguous().numpy(), g[f"aug{i}"]), v.name
        assert np.array_equal(v.deaugment_mask(a).contiguous().numpy(), g[f"rt{i}"])
        assert torch.equal(v.deaugment_mask(a), xt)
    assert len(oinf.flip8_tta()) == 8


@pytest.mark.parametrize("tag,use_gc", [("nogc", False), ("gc", True)])
def test_ranger_matches_reference(golden_dir, tag, use_gc):
    g = _load(golden_dir, "ranger.npz")
    gen = torch.Generator().manual_seed(11)
    p0 = [torch.randn(6, 5, 3, 3, 3, generator=gen), torch.randn(7, generator=gen), torch.randn(4, 6, generator=gen)]
    grads = [[torch.randn(p.shape, generator=gen) * 0.1 for p in p0] for _ in range(14)]
    ps = [p.clone() for p in p0]
    states = [train.RangerState(p) for p in ps]
    for step in range(14):
        train.ranger_step(ps, grads[step], states, lr=3e-4, weight_decay=1e-5, use_gc=use_gc)
        if step in (4, 5, 13):
            for i, p in enumerate(ps):
                ref = torch.from_numpy(g[f"{tag}_s{step + 1}_p{i}"])
                assert (p - ref).abs().max().item() <= 1e-6, (tag, step, i)


def test_dice_known_values():
    # MONAI DiceLoss is not vendored (parity unpinned): known-answer checks of the restated formula only.
    t = synth.target((16, 16, 16))
    big = (t * 2 - 1) * 50.0
    assert train.dice_loss(big, t).item() < 1e-6
    assert abs(train.dice_loss(-big, t).item() - 1.0) < 1e-4
    z = torch.zeros_like(t)
    # p = 0.5 everywhere: I = 0.5*sum(t), P = 0.25*V, G = sum(t)
    s, v = t.sum(dim=(0, 2, 3, 4)), float(16 ** 3)
    exp = (1 - (s + 1e-5) / (s + 0.25 * v + 1e-5)).mean().item()
    assert abs(train.dice_loss(z, t).item() - exp) < 1e-6
    expj = (1 - (s + 1e-5) / (2 * (s + 0.25 * v - 0.5 * s) + 1e-5)).mean().item()
    assert abs(train.dice_loss(z, t, jaccard=True).item() - expj) < 1e-6


def test_label_postprocessing():
    oh = torch.zeros(1, 3, 2, 2, 2)
    oh[0, 1] = 1            # WT everywhere
    oh[0, 0, 0] = 1         # TC on the first plane
    oh[0, 2, 0, 0] = 1      # ET on the first row of it
    lab = oinf.brats_label_map(oh)
    assert lab.dtype == torch.uint8 and lab.shape == (1, 1, 2, 2, 2)
    assert lab[0, 0, 0, 0].tolist() == [4, 4] and lab[0, 0, 0, 1].tolist() == [1, 1] and (lab[0, 0, 1] == 2).all()
    img = torch.zeros(1, 4, 2, 2, 2)
    img[0, 1, 0] = -0.5
    out = oinf.remove_background_voxels(img, torch.ones(1, 3, 2, 2, 2))
    assert out[0, :, 0].min() == 1 and out[0, :, 1].max() == 0
    x = torch.zeros(1, 4, 240, 240, 155)
    y, pb, pa = oinf.shape_to_divisible(x, 8)
    assert y.shape[-3:] == (240, 240, 160) and pb.tolist() == [0, 0, 3] and pa.tolist() == [0, 0, 2]
    assert oinf.shape_to_original(y, pb, pa).shape == x.shape


def _dice_ce_inputs():
    g = torch.Generator().manual_seed(23)
    x = torch.randn((2, 3, 6, 5, 7), generator=g) * 2.0
    wt = torch.rand((2, 1, 6, 5, 7), generator=g) > 0.5
    tc = wt & (torch.rand((2, 1, 6, 5, 7), generator=g) > 0.4)
    et = tc & (torch.rand((2, 1, 6, 5, 7), generator=g) > 0.5)
    return x, torch.cat([tc, wt, et], dim=1).float()


@pytest.mark.parametrize("tag,kw", [("default", {}), ("weighted", {"lambda_dice": 0.7, "lambda_ce": 1.3})])
def test_dice_ce_loss_matches_reference(golden_dir, tag, kw):
    """oracle.train.dice_ce_loss against the unmodified learning.losses.DiceCELoss (tests/golden/make_golden_losses.py)."""
    g = _load(golden_dir, "dice_ce.npz")
    x, t = _dice_ce_inputs()
    xr = x.clone().requires_grad_(True)
    loss = train.dice_ce_loss(xr, t, **kw)
    loss.backward()
    assert abs(loss.item() - float(g[f"{tag}_loss"])) <= 1e-6
    assert np.abs(xr.grad.numpy() - g[f"{tag}_grad"]).max() <= 1e-7


V1_VARIANTS = (("instance", "leakyrelu"), ("batch", "elu"), ("none", "relu"), ("group", "leakyrelu"), ("instance", "relu"))


@pytest.mark.parametrize("norm,act", V1_VARIANTS)
def test_v1_norm_act_factory_matches_reference(golden_dir, norm, act):
    """EquiUnet under the other norms / activations of networks/factory.py:179-200 (unmodified reference golden)."""
    g = _load(golden_dir, "net_v1_w16_variants.npz")
    tag = f"{norm}_{act}"
    x = synth.volume(seed=3, shape=(16, 16, 16))
    p = synth.make_params(1, WIDTH, 123, norm=norm)
    assert list(p.keys()) == list(g[f"{tag}_keys"])  # state_dict key ORDER too
    with torch.no_grad():
        out, deeps = nets.equiunet_v1_forward(p, x, norm=norm, act=act)
    ref = torch.from_numpy(g[f"{tag}_out"])
    assert (out - ref).abs().max().item() <= 2e-4 * ref.abs().max().item()
    r0 = torch.from_numpy(g[f"{tag}_deep0_s2"])
    assert (deeps[0][..., ::2, ::2, ::2] - r0).abs().max().item() <= 2e-4 * max(r0.abs().max().item(), 1.0)
    if norm == "batch":
        with torch.no_grad():
            out, _ = nets.equiunet_v1_forward(p, x, norm=norm, act=act, training=True)
        ref = torch.from_numpy(g[f"{tag}_train_out"])
        assert (out - ref).abs().max().item() <= 2e-4 * ref.abs().max().item()
        assert np.abs(p["encoder1.ConvBnRelu1.bn.running_mean"].numpy() - g[f"{tag}_train_running_mean"]).max() <= 1e-5
        assert np.abs(p["decoder1.ConvBnRelu2.bn.running_var"].numpy() - g[f"{tag}_train_running_var"]).max() <= 1e-5
