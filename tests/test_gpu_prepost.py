"""GPU parity of the pre-/post-processing kernels (csrc/pre.cu, csrc/post.cu) through the C-ABI against the oracle
(oracle/prepost.py) and the reference golden vectors (tests/golden/prepost.npz).  Integer work is bit-exact;
normalisation is float: |err| <= 2e-5 (fp64 device sums vs numpy's fp32 pairwise mean/std)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "prepost.npz"))


@pytest.mark.parametrize("i,seed", [(0, 0), (1, 1)])
@pytest.mark.parametrize("ro", [False, True])
def test_normalize_pad_matches_reference_golden(gold, i, seed, ro):
    from brats21_b200 import preprocess
    from oracle import prepost as pp
    img = torch.from_numpy(pp.synth_raw(seed)).to(DEV)
    out, meta = preprocess.crop_normalize_pad(img, 8, remove_outliers=ro, crop_foreground=False)
    ref = gold[f"norm{i}_ro{int(ro)}"]
    assert tuple(out.shape[1:]) == ref.shape
    assert list(meta.pad_before) == list(gold[f"norm{i}_pb"]) and list(meta.pad_after) == list(gold[f"norm{i}_pa"])
    got = out[0].cpu().numpy()
    assert np.array_equal(got == 0, ref == 0)
    assert np.abs(got - ref).max() <= 2e-5


@pytest.mark.parametrize("shape", [(4, 20, 22, 19), (4, 37, 41, 33), (1, 9, 8, 70)])
def test_crop_normalize_pad_matches_oracle(shape):
    from brats21_b200 import preprocess
    from oracle import prepost as pp
    raw = pp.synth_raw(7, shape)
    out, meta = preprocess.crop_normalize_pad(torch.from_numpy(raw).to(DEV), 8, remove_outliers=True)
    ref, start, end, pb, pa = pp.preprocess(raw, 8, remove_outliers=True)
    assert list(meta.start) == start and list(meta.end) == end
    assert list(meta.pad_before) == list(pb) and list(meta.pad_after) == list(pa)
    got = out[0].cpu().numpy()
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() <= 2e-5
    # stats entry point on its own: count is exact, sums to fp64 round-off
    bbox = preprocess.foreground_bbox(torch.from_numpy(raw).to(DEV))
    st = preprocess.nonzero_stats(torch.from_numpy(raw).to(DEV), bbox).cpu().numpy()
    for c in range(shape[0]):
        v = raw[c][raw[c] != 0].astype(np.float64)
        assert st[c, 0] == v.size
        assert abs(st[c, 1] - v.sum()) <= 1e-6 * abs(v.sum()) and abs(st[c, 2] - (v * v).sum()) <= 1e-6 * (v * v).sum()


def test_preprocess_edge_cases():
    from brats21_b200 import preprocess
    z = torch.zeros((4, 8, 9, 10), device=DEV)
    out, meta = preprocess.crop_normalize_pad(z)  # no foreground: whole volume, all zeros
    assert out.shape == (1, 4, 8, 16, 16) and out.abs().max().item() == 0 and meta.start == (0, 0, 0)
    c = torch.zeros((2, 8, 8, 8), device=DEV)
    c[0, 2:5, 3, 4] = 7.0  # constant channel: std 0 -> 1, (x - mean) = 0
    c[1, 2:5, 3, 4] = torch.tensor([1.0, 2.0, 3.0], device=DEV)
    out, meta = preprocess.crop_normalize_pad(c)
    assert meta.start == (2, 3, 4) and meta.end == (5, 4, 5) and out.shape == (1, 2, 8, 8, 8)
    core = out[0, :, 3:6, 4, 4]  # ceil(5/2) = 3 before on d, ceil(7/2) = 4 on h and w
    assert core[0].abs().max().item() == 0
    assert torch.allclose(core[1].cpu(), torch.tensor([-1.2247449, 0.0, 1.2247449]), atol=1e-6)
    with pytest.raises(RuntimeError):
        preprocess.crop_normalize_pad(torch.zeros((4, 8, 8, 8)))  # CPU tensor: no fallback


@pytest.mark.parametrize("i,seed", [(0, 3), (1, 4)])
def test_keep_components_matches_reference_golden(gold, i, seed):
    from brats21_b200.postprocess import KeepLargestConnectedComponent
    from oracle import prepost as pp
    lab = torch.from_numpy(pp.synth_labels(seed)).to(DEV)[None, None]
    for thr in (None, 1, 10, 12, 100000):
        out = KeepLargestConnectedComponent(thr)(lab)
        assert out.dtype == torch.uint8 and np.array_equal(out[0, 0].cpu().numpy(), gold[f"cc{i}_t{thr}"]), thr
    f = KeepLargestConnectedComponent(10)(lab.float())  # the reference hands float tensors around
    assert f.dtype == torch.float32 and np.array_equal(f[0, 0].cpu().numpy().astype(np.uint8), gold[f"cc{i}_t10"])


@pytest.mark.parametrize("density", [0.02, 0.2, 0.5])
def test_keep_components_random_masks_match_oracle(density):
    """Random masks near the percolation threshold: long, winding components stress the union-find."""
    from brats21_b200.postprocess import KeepLargestConnectedComponent
    from oracle import prepost as pp
    g = np.random.default_rng(int(density * 100))
    lab = ((g.random((40, 37, 45)) < density) * g.integers(1, 5, size=(40, 37, 45))).astype(np.uint8)
    t = torch.from_numpy(lab).to(DEV)[None, None]
    for thr in (None, 0, 3, 50):
        ref = pp.get_largest_component(lab, thr)
        got = KeepLargestConnectedComponent(thr)(t)[0, 0].cpu().numpy()
        if thr is None:
            # equal-size largest components: first in raster order (np.argmax) — same rule on both sides
            pass
        assert np.array_equal(got, ref), (density, thr)


def test_keep_components_full_size_properties():
    """BASELINE-size volume (240 x 240 x 155): idempotence, monotonicity in the threshold, empty input."""
    from brats21_b200.postprocess import KeepLargestConnectedComponent
    g = torch.Generator(device="cpu").manual_seed(0)
    lab = (torch.rand((240, 240, 155), generator=g) < 0.3).to(torch.uint8).to(DEV)[None, None] * 2
    a = KeepLargestConnectedComponent(10)(lab)
    assert torch.equal(KeepLargestConnectedComponent(10)(a), a)
    b = KeepLargestConnectedComponent(1000)(lab)
    assert torch.all((b != 0) <= (a != 0)) and torch.all((a != 0) <= (lab != 0))
    big = KeepLargestConnectedComponent(None)(lab)
    assert torch.equal(KeepLargestConnectedComponent(None)(big), big) and torch.all((big != 0) <= (b != 0))
    z = torch.zeros_like(lab)
    assert torch.equal(KeepLargestConnectedComponent(None)(z), z) and torch.equal(KeepLargestConnectedComponent(5)(z), z)


@pytest.mark.parametrize("i,seed", [(0, 3), (1, 4)])
@pytest.mark.parametrize("axis", [0, 1, 2])
def test_replace_rare_labels_matches_reference_golden(gold, i, seed, axis):
    from brats21_b200.postprocess import ReplaceWithClosestValue
    from oracle import prepost as pp
    lab = pp.synth_labels(seed)
    ref_o, amb = pp.replace_with_closest_value(lab, 20, axis)
    got = ReplaceWithClosestValue(labels=[3], thresh=20, axis=axis)(torch.from_numpy(lab).to(DEV)[None, None])
    got = got[0, 0].cpu().numpy()
    assert np.array_equal(got, ref_o)  # oracle incl. its documented tie rule: bit-exact
    ref = gold[f"rep{i}_a{axis}"]
    assert np.array_equal(got[~amb], ref[~amb])  # unmodified reference wherever griddata's answer is unique


def test_replace_rare_labels_noop_and_channels():
    from brats21_b200.postprocess import ReplaceWithClosestValue, labels_to_channels
    from oracle import prepost as pp
    lab = pp.synth_labels(3)
    t = torch.from_numpy(lab).to(DEV)[None, None]
    same = ReplaceWithClosestValue(thresh=2)(t)  # no label has <= 2 voxels: untouched
    assert torch.equal(same, t)
    ch = labels_to_channels(t)
    assert np.array_equal(ch[0].cpu().numpy(), pp.labels_to_channels(lab))
    with pytest.raises(AssertionError):
        ReplaceWithClosestValue()(t[0])


def test_ranger_gradient_centralisation_matches_reference_golden(golden_dir):
    """use_gc=True path of Ranger2020 (learning/optimizer.py:11-20,187-188) against the reference golden."""
    from brats21_b200.optimizer import Ranger2020
    g = np.load(os.path.join(golden_dir, "ranger.npz"))
    gen = torch.Generator().manual_seed(11)
    p0 = [torch.randn(6, 5, 3, 3, 3, generator=gen), torch.randn(7, generator=gen), torch.randn(4, 6, generator=gen)]
    grads = [[torch.randn(p.shape, generator=gen) * 0.1 for p in p0] for _ in range(14)]
    ps = [torch.nn.Parameter(p.clone().to(DEV)) for p in p0]
    opt = Ranger2020(ps, lr=3e-4, alpha=0.5, k=6, N_sma_threshhold=5, betas=(.95, 0.999), eps=1e-5, weight_decay=1e-5,
                     use_gc=True, gc_conv_only=False)
    for p in ps:
        p.grad = torch.zeros_like(p)
    for step in range(14):
        for p, gr in zip(ps, grads[step]):
            p.grad.copy_(gr)
        opt.step()
        if step in (4, 5, 13):
            for i, p in enumerate(ps):
                ref = torch.from_numpy(g[f"gc_s{step + 1}_p{i}"])
                assert (p.detach().cpu() - ref).abs().max().item() <= 2e-5 * max(ref.abs().max().item(), 1.0)


def test_predict_case_end_to_end_matches_oracle_chain():
    """Raw intensities -> labels at the original size through engine.predict_case, against the oracle chain
    (crop/normalise/pad -> V2 sliding window with 8 flips -> mean/threshold/background -> component filter -> pad
    back).  Labels are compared outside the 0.02 probability margin (bf16 network vs fp32 oracle)."""
    from brats21_b200 import engine, networks, tta
    from oracle import inference as oinf
    from oracle import nets, synth
    from oracle import prepost as pp
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    width = 16
    params = {k: v.to(DEV) for k, v in synth.make_params(2, width, 93).items()}
    net = networks.EquiUnetASSPEvo(4, 3, [width * 2 ** i for i in range(4)], deep_supervision=True).to(DEV).eval()
    net.load_state_dict(params)
    raw = pp.synth_raw(5, (4, 50, 44, 41))
    got = engine.predict_case([net], torch.from_numpy(raw).to(DEV), tta.get_flip8_transforms(), (32, 32, 32), 2,
                              cleaning_areas_threshold=10)
    assert got.shape == raw.shape[1:] and got.dtype == torch.uint8
    vol, start, end, pb, pa = pp.preprocess(raw, 8)
    v = torch.from_numpy(vol)[None]
    with torch.no_grad():
        fwd = lambda z: nets.equiunet_v2_forward(params, z.to(DEV))[0].cpu()  # noqa: E731
        outs = oinf.apply_tta(lambda z: oinf.sliding_window_inference(z.contiguous(), (32, 32, 32), 2, fwd, 0.25,
                                                                      "gaussian"), v, oinf.flip8_tta())
        prob, hard = oinf.ensemble_mean_threshold(outs)
        hard = oinf.remove_background_voxels(v, hard)
    margin_ok = ((prob - 0.5).abs() > 0.02).all(dim=1)[0].numpy()
    lab = pp.brats_label_map(hard[0].numpy())
    # voxels inside the margin may legitimately differ and, through them, so may component sizes: compare the
    # un-cleaned maps voxel-wise and the cleaned ones only when nothing sits inside the margin of a small component
    plain = engine.predict_case([net], torch.from_numpy(raw).to(DEV), tta.get_flip8_transforms(), (32, 32, 32), 2)
    core = lab[pb[0]:lab.shape[0] - pa[0], pb[1]:lab.shape[1] - pa[1], pb[2]:lab.shape[2] - pa[2]]
    full = np.zeros(raw.shape[1:], dtype=np.uint8)
    full[start[0]:end[0], start[1]:end[1], start[2]:end[2]] = core
    mfull = np.ones(raw.shape[1:], dtype=bool)
    mfull[start[0]:end[0], start[1]:end[1], start[2]:end[2]] = \
        margin_ok[pb[0]:lab.shape[0] - pa[0], pb[1]:lab.shape[1] - pa[1], pb[2]:lab.shape[2] - pa[2]]
    p = plain.cpu().numpy()
    assert np.array_equal(p[mfull], full[mfull])
    assert p[~np.any(raw != 0, axis=0)].max() == 0  # background removed, zero outside the crop box
    # reference order (learning/engine.py:249-256): threshold -> post transforms -> remove_background_voxels.  The
    # cleaned result equals the oracle's component filter applied to OUR un-masked label map, then the background
    # mask, then the pad-back (integer stages: exact)
    from brats21_b200 import preprocess
    vol_d, meta = preprocess.crop_normalize_pad(torch.from_numpy(raw).to(DEV), 8)
    _, nobg = engine.predict_volume([net], vol_d, tta.get_flip8_transforms(), True, (32, 32, 32), 2,
                                    remove_background=False)
    cleaned = pp.get_largest_component(nobg[0, 0].cpu().numpy(), 10)
    cleaned[~np.any(vol_d[0].cpu().numpy() != 0, axis=0)] = 0
    want = np.zeros(raw.shape[1:], dtype=np.uint8)
    want[start[0]:end[0], start[1]:end[1], start[2]:end[2]] = \
        cleaned[pb[0]:lab.shape[0] - pa[0], pb[1]:lab.shape[1] - pa[1], pb[2]:lab.shape[2] - pa[2]]
    assert np.array_equal(got.cpu().numpy(), want)


def test_post_transforms_factory_with_cleaning():
    import argparse
    from brats21_b200 import definer
    from oracle import prepost as pp
    lab = pp.synth_labels(4)
    onehot = pp.labels_to_channels(lab).astype(np.float32)
    prob = torch.from_numpy(onehot * 0.9 + 0.05)[None].to(DEV)
    args = argparse.Namespace(cleaning_areas=True, cleaning_areas_threshold=10, replace_value=True,
                              replace_value_threshold=20)
    out = definer.get_post_transforms(args)(prob)
    ref = pp.get_largest_component(lab, 10)
    ref, amb = pp.replace_with_closest_value(ref, 20, 2)
    assert out.shape == (1, 3) + lab.shape and out.dtype == torch.float32
    assert np.array_equal(out[0].cpu().numpy().astype(np.uint8), pp.labels_to_channels(ref))
    plain = definer.get_post_transforms(argparse.Namespace())(prob)
    assert np.array_equal(plain[0].cpu().numpy(), onehot)
