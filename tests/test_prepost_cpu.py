"""CPU: the pre-/post-processing oracle (oracle/prepost.py) against tests/golden/prepost.npz, produced by the
unmodified reference utils/transforms.py (tests/golden/make_golden_prepost.py)."""
import os

import numpy as np
import pytest

from oracle import prepost as pp


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "prepost.npz"))


@pytest.mark.parametrize("i,seed", [(0, 0), (1, 1)])
@pytest.mark.parametrize("ro", [False, True])
def test_normalize_and_pad_match_reference(gold, i, seed, ro):
    img = pp.synth_raw(seed)
    out, pb, pa = pp.shape_to_divisible(pp.normalize_intensity(img, remove_outliers=ro), 8)
    assert list(pb) == list(gold[f"norm{i}_pb"]) and list(pa) == list(gold[f"norm{i}_pa"])
    ref = gold[f"norm{i}_ro{int(ro)}"]
    assert out.shape == ref.shape
    np.testing.assert_allclose(out, ref, rtol=0, atol=1e-6)  # same numpy ops: float32 round-off only


@pytest.mark.parametrize("i,seed", [(0, 3), (1, 4)])
def test_components_match_reference(gold, i, seed):
    lab = pp.synth_labels(seed)
    for thr in (None, 1, 10, 12, 100000):
        out = pp.get_largest_component(lab, thr)
        assert np.array_equal(out, gold[f"cc{i}_t{thr}"]), thr
    # threshold semantics: strictly more than `threshold` voxels survive (12-voxel islands go at 12, stay at 10)
    assert (gold[f"cc{i}_t10"] != 0).sum() > (gold[f"cc{i}_t12"] != 0).sum()
    assert (gold[f"cc{i}_t100000"] != 0).sum() == 0


@pytest.mark.parametrize("i,seed", [(0, 3), (1, 4)])
@pytest.mark.parametrize("axis", [0, 1, 2])
def test_replace_rare_matches_reference_where_unambiguous(gold, i, seed, axis):
    lab = pp.synth_labels(seed)
    out, amb = pp.replace_with_closest_value(lab, 20, axis)
    ref = gold[f"rep{i}_a{axis}"]
    assert (lab == 4).sum() > 0 and (ref == 4).sum() == 0  # the rare ET voxels were replaced
    assert np.array_equal(out[~amb], ref[~amb])
    # where the KD-tree had to choose between equidistant neighbours, the reference value is one of the candidates
    assert amb.sum() < (lab == 4).sum() + 1


@pytest.mark.parametrize("i,seed", [(0, 3), (1, 4)])
def test_label_map_round_trip(gold, i, seed):
    lab = pp.synth_labels(seed)
    onehot = np.stack([(lab == 1) | (lab == 4), lab > 0, lab == 4]).astype(np.uint8)
    assert np.array_equal(pp.brats_label_map(onehot), gold[f"brats{i}"])
    assert np.array_equal(pp.labels_to_channels(gold[f"brats{i}"]), onehot)


def test_preprocess_crop_bbox_and_zero_background():
    img = pp.synth_raw(0)
    out, start, end, pb, pa = pp.preprocess(img)
    assert all(s % 8 == 0 for s in out.shape[1:])
    fg = np.any(img > 0, axis=0)
    assert fg[start[0]:end[0], start[1]:end[1], start[2]:end[2]].sum() == fg.sum()
    core = out[:, pb[0]:out.shape[1] - pa[0], pb[1]:out.shape[2] - pa[1], pb[2]:out.shape[3] - pa[2]]
    crop = img[:, start[0]:end[0], start[1]:end[1], start[2]:end[2]]
    assert np.array_equal(core == 0, crop == 0) or np.abs(core[(core == 0) != (crop == 0)]).max() == 0
    for c in range(img.shape[0]):
        v = core[c][crop[c] != 0]
        assert abs(v.mean()) < 1e-4 and abs(v.std() - 1) < 1e-4


def test_pad_back_inverts_crop_and_pad():
    """postprocess.pad_back (pure slicing, device-agnostic) undoes oracle.preprocess's crop + pad bookkeeping."""
    import torch
    from brats21_b200.postprocess import pad_back
    from brats21_b200.preprocess import CropMeta
    img = pp.synth_raw(2, (2, 21, 18, 27))
    out, start, end, pb, pa = pp.preprocess(img, 8)
    meta = CropMeta(tuple(img.shape[1:]), tuple(start), tuple(end), tuple(int(v) for v in pb), tuple(int(v) for v in pa))
    marker = torch.arange(out[0].size, dtype=torch.float32).reshape(out.shape[1:]) + 1.0
    back = pad_back(marker, meta)
    assert tuple(back.shape) == img.shape[1:]
    core = marker[pb[0]:marker.shape[0] - pa[0], pb[1]:marker.shape[1] - pa[1], pb[2]:marker.shape[2] - pa[2]]
    assert torch.equal(back[start[0]:end[0], start[1]:end[1], start[2]:end[2]], core)
    outside = back.clone()
    outside[start[0]:end[0], start[1]:end[1], start[2]:end[2]] = 0
    assert outside.abs().max().item() == 0
