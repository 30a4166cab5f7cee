"""GPU parity of the sliding-window / TTA / label pipeline against the reference's golden vectors and the oracle."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(autouse=True)
def _no_tf32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


def _ramp_predictor(x):
    d, h, w = x.shape[2:]
    i = torch.arange(d, dtype=x.dtype, device=x.device).reshape(1, d, 1, 1)
    j = torch.arange(h, dtype=x.dtype, device=x.device).reshape(1, 1, h, 1)
    k = torch.arange(w, dtype=x.dtype, device=x.device).reshape(1, 1, 1, w)
    ramp = 1.0 + 0.01 * (i + 2 * j + 3 * k)
    y = torch.stack([x[:, 0] * 0.5 + x[:, 1], x[:, 2] - x[:, 3], x.sum(1) * 0.25], dim=1)
    return y * ramp.unsqueeze(0)


def test_sliding_window_generic_predictor_matches_reference_golden(golden_dir):
    """Same predictor as tests/golden/make_golden.py ran through the unmodified utils/inferers.py."""
    from brats21_b200.inferers import sliding_window_inference
    from oracle import synth
    g = np.load(os.path.join(golden_dir, "sliding_window.npz"))
    xs = synth.volume(seed=5, shape=(40, 36, 29)).to(DEV)
    for mode in ("constant", "gaussian"):
        for bs in (1, 4):
            y = sliding_window_inference(xs, (16, 16, 16), bs, _ramp_predictor, overlap=0.25, mode=mode)
            ref = torch.from_numpy(g[f"{mode}_b{bs}"])
            assert y.shape == ref.shape and y.dtype == torch.float32
            assert (y.cpu() - ref).abs().max().item() <= 2e-6 * ref.abs().max().item()
    y = sliding_window_inference(xs, (48, 32, 16), 2, _ramp_predictor, overlap=0.5, mode="gaussian",
                                 device=torch.device("cpu"))
    assert y.device.type == "cpu"
    ref = torch.from_numpy(g["pad_gaussian"])
    assert (y - ref).abs().max().item() <= 2e-6 * ref.abs().max().item()
    with pytest.raises(AssertionError):
        sliding_window_inference(xs, 16, 1, _ramp_predictor, overlap=1.0)


def _net(width=16, seed=93):
    import warnings
    from brats21_b200 import networks
    from oracle import synth
    params = {k: v.to(DEV) for k, v in synth.make_params(2, width, seed).items()}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        net = networks.EquiUnetASSPEvo(4, 3, [width * 2 ** i for i in range(4)], deep_supervision=True).to(DEV).eval()
    net.load_state_dict(params)
    return net, params


def test_sliding_window_fast_path_matches_oracle():
    from brats21_b200.inferers import sliding_window_inference
    from oracle import inference as oinf
    from oracle import nets, synth
    net, params = _net()
    vol = torch.cat([synth.volume(seed=s, shape=(48, 40, 24)) for s in (1, 2)]).to(DEV)
    fwd = lambda z: nets.equiunet_v2_forward(params, z.to(DEV))[0].cpu()  # noqa: E731
    for mode, bs in (("constant", 1), ("gaussian", 3)):
        with torch.no_grad():
            ref = oinf.sliding_window_inference(vol.cpu(), (32, 32, 32), bs, fwd, 0.25, mode)
        got = sliding_window_inference(vol, (32, 32, 32), bs, net, overlap=0.25, mode=mode)
        assert got.shape == ref.shape
        rel = ((got.cpu() - ref).norm() / ref.norm()).item()
        assert rel <= 2.5e-2, rel
        # fast path vs generic path through the same module: same kernels, same blend order; the only difference
        # is the summation order of the atomically accumulated norm statistics (last-bit effects after bf16 rounding)
        gen = sliding_window_inference(vol, (32, 32, 32), bs, lambda z: net(z), overlap=0.25, mode=mode)
        assert ((got - gen).norm() / gen.norm()).item() <= 1.5e-2


@pytest.mark.parametrize("which", ["tta16", "flip8", "none_full"])
def test_predict_volume_matches_oracle_pipeline(which):
    from brats21_b200 import engine, tta
    from oracle import inference as oinf
    from oracle import nets, synth
    net, params = _net()
    net2, params2 = _net(seed=7)
    vol = synth.volume(seed=11, shape=(40, 32, 48)).to(DEV)
    roi = (32, 32, 32)
    sw = which != "none_full"
    comp, ovar = {"tta16": (tta.get_tta_transforms(), oinf.reference_tta()),
                  "flip8": (tta.get_flip8_transforms(), oinf.flip8_tta()),
                  "none_full": (None, [oinf.Variant("id", lambda x: x, lambda x: x)])}[which]
    outs = []
    with torch.no_grad():
        for p in (params, params2):
            fwd = lambda z, p=p: nets.equiunet_v2_forward(p, z.contiguous().to(DEV))[0].cpu()  # noqa: E731
            f = (lambda z: oinf.sliding_window_inference(z.contiguous().cpu(), roi, 2, fwd, 0.25, "gaussian")) if sw else fwd
            outs += oinf.apply_tta(f, vol.cpu(), ovar)
    prob_ref, hard_ref = oinf.ensemble_mean_threshold(outs)
    hard_ref = oinf.remove_background_voxels(vol.cpu(), hard_ref)
    lab_ref = oinf.brats_label_map(hard_ref)
    onehot, label, prob = engine.predict_volume([net, net2], vol, comp, sw, roi, 2, 0.25, "gaussian", return_prob=True)
    assert onehot.shape == (1, 3, 40, 32, 48) and onehot.dtype == torch.uint8 and label.shape == (1, 1, 40, 32, 48)
    tol = 0.02  # probability tolerance implied by the bf16 logit tolerance
    assert (prob.cpu()[None] - prob_ref).abs().max().item() <= tol
    margin = (prob_ref - 0.5).abs() > tol
    assert not ((onehot.cpu().float() != hard_ref) & margin).any()  # bit-exact labels outside the margin
    agree = (label.cpu() == lab_ref).float().mean().item()
    assert agree >= 0.999, agree
    # Per-region Dice agreement (TC, WT, ET).  Outside the margin the maps are bit-exact (Dice 1.0, asserted above).
    # Over ALL voxels the bound is set by the density of random-init logits near the threshold, not by the kernels:
    # torch's own bf16 autocast reaches only 0.992-0.999 on these networks, so 0.995 is asserted
    # for regions large enough for the ratio to be meaningful.
    for c in range(3):
        a, b = onehot[0, c].cpu().bool(), hard_ref[0, c].bool()
        if b.sum() >= 5000:
            assert 2 * (a & b).sum().item() / (a.sum().item() + b.sum().item()) >= 0.995
    assert (label.cpu()[0, 0][(vol.cpu()[0] == 0).all(0)] == 0).all()  # background removed


def test_full_size_properties():
    """BASELINE-size (240x240x160) checks that need no oracle run: partition of unity and TTA round trip."""
    from brats21_b200 import ops, tta
    from brats21_b200.inferers import WindowPlan, sliding_window_inference
    vol = torch.randn((1, 4, 240, 240, 160), device=DEV)
    # a predictor that returns (three of) its input channels: blending must reproduce the volume itself
    for mode in ("constant", "gaussian"):
        y = sliding_window_inference(vol, (128, 128, 128), 4, lambda z: z[:, :3], overlap=0.25, mode=mode)
        assert (y - vol[:, :3]).abs().max().item() <= 1e-5
    plan = WindowPlan((240, 240, 160), (128,) * 3, 0.25, "gaussian", 0.125, vol.device)
    assert len(plan.origins) == 18 and plan.count.min().item() > 0
    # pack -> (identity network) -> blend -> de-augment returns the source volume for every variant
    for tr in list(tta.get_tta_transforms())[::5] + list(tta.get_flip8_transforms())[::3]:
        perm, flip = tr.variant
        adims = [0, 0, 0]
        for j in range(3):
            adims[perm[j]] = vol.shape[2 + j]
        p = WindowPlan(adims, (128,) * 3, 0.25, "gaussian", 0.125, vol.device)
        acc = torch.zeros((3,) + p.image_size, device=DEV)
        for g0 in range(0, len(p.origins), 4):
            grp = p.origins[g0:g0 + 4]
            x8 = torch.empty((len(grp), 128, 128, 128, 8), device=DEV, dtype=torch.bfloat16)
            ops.pack_windows(vol, x8, grp, perm=perm, flip=flip)
            logits = x8[..., :3].float().permute(0, 4, 1, 2, 3).contiguous()
            ops.blend_accumulate(logits, acc, p.profiles, grp)
        out = torch.empty((3, 240, 240, 160), device=DEV)
        ops.tta_accumulate(acc, p.count, out, perm, flip, apply_sigmoid=False, overwrite=True)
        assert (out - vol[0, :3].to(torch.bfloat16).float()).abs().max().item() <= 1e-5
