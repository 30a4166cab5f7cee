"""TEST INFRASTRUCTURE ONLY — CPU restatement (numpy / scipy) of the reference's pre- and post-processing around the
hot path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import it.

Pinned against the UNMODIFIED reference functions (utils/transforms.py: NormalizeIntensity, shape_to_divisible,
get_largest_component, replace_w_closest_value_3d, ConvertToBratsClassesBasedOnMultiChannel) run in the build
container by tests/golden/make_golden_prepost.py -> tests/golden/prepost.npz.  Un-vendored third-party pieces,
restated from their published semantics — parity unpinned for those:
  * MONAI 0.6.0 ``CropForeground`` (``generate_spatial_bounding_box``: select_fn = x > 0 over any channel, margin 0)
    and ``ConvertToMultiChannelBasedOnBratsClasses`` (TC = 1|4, WT = 1|2|4, ET = 4);
  * skimage ``morphology.label`` (default connectivity = ndim, i.e. 26 neighbours in 3-D, labels numbered in raster
    order), restated with ``scipy.ndimage.label(structure=ones(3,3,3))``;
  * scipy ``griddata(method="nearest")`` is present in this image and used directly where its answer is unique.
"""
from __future__ import annotations

import numpy as np


def crop_foreground_bbox(img: np.ndarray):
    """MONAI generate_spatial_bounding_box(img, select_fn=lambda x: x > 0, channel_indices=None, margin=0)."""
    fg = np.any(img > 0, axis=0)
    nz = np.nonzero(fg)
    start = [int(a.min()) for a in nz]
    end = [int(a.max()) + 1 for a in nz]
    return start, end


def normalize_intensity(img: np.ndarray, remove_outliers: bool = False, outliers_value: float = 3.0) -> np.ndarray:
    """NormalizeIntensity(nonzero=True, channel_wise=True) (utils/transforms.py:363-406)."""
    out = img.astype(np.float32).copy()
    for c in range(out.shape[0]):
        ch = out[c]
        sl = ch != 0
        if not np.any(sl):
            continue
        sub = np.mean(ch[sl])
        div = np.std(ch[sl])
        if div == 0.0:
            div = 1.0
        ch[sl] = (ch[sl] - sub) / div
        if remove_outliers:
            ch[sl] = np.clip(ch[sl], -outliers_value, outliers_value)
    return out


def shape_to_divisible(data: np.ndarray, k: int = 8):
    """utils/transforms.py:483-512 on the last three dims: ceil(p/2) before, floor(p/2) after."""
    shape = np.array(data.shape[-3:])
    target = np.ceil(shape / k).astype(int) * k
    p = target - shape
    pb = np.ceil(p / 2).astype(int)
    pa = np.floor(p / 2).astype(int)
    pads = [(0, 0)] * (data.ndim - 3) + [(int(pb[i]), int(pa[i])) for i in range(3)]
    return np.pad(data, pads), pb, pa


def preprocess(img: np.ndarray, k: int = 8, remove_outliers: bool = False):
    """CropForegroundd -> NormalizeIntensityd -> shape_to_divisible (src/definer.py:561-567, engine.py:229)."""
    start, end = crop_foreground_bbox(img)
    crop = img[:, start[0]:end[0], start[1]:end[1], start[2]:end[2]]
    norm = normalize_intensity(crop, remove_outliers)
    out, pb, pa = shape_to_divisible(norm, k)
    return out, start, end, pb, pa


def get_largest_component(in_volume: np.ndarray, threshold=None) -> np.ndarray:
    """utils/transforms.py:579-600 with skimage.morphology.label restated through scipy.ndimage.label."""
    from scipy import ndimage
    vol = in_volume.copy()
    mask = vol != 0
    lbls, n = ndimage.label(mask, structure=np.ones((3,) * mask.ndim, dtype=bool))
    if n == 0:
        return vol
    sizes = np.bincount(lbls.ravel(), minlength=n + 1)
    if threshold is None:
        region = np.array([np.argmax(sizes[1:]) + 1])
    else:
        region = np.nonzero(sizes[1:] > threshold)[0] + 1
    vol[~np.isin(lbls, region)] = 0
    return vol


def replace_with_closest_value(label: np.ndarray, thresh: int = 20, axis: int = 2):
    """ReplaceWithClosestValue.__call__ + replace_w_closest_value_3d (utils/transforms.py:254-268, 603-647) on a
    [D0, D1, D2] array.  Returns (result, ambiguous): griddata's choice between EQUIDISTANT nearest neighbours with
    different values is an implementation detail of the KD-tree; such voxels are flagged in ``ambiguous`` and the
    result holds the smallest-row-major-index candidate (the rule the CUDA kernel documents)."""
    arr = label.copy()
    amb = np.zeros(arr.shape, dtype=bool)
    uniq, counts = np.unique(arr, return_counts=True)
    values = uniq[counts <= thresh]
    if not values.any():
        return arr, amb
    out = arr.copy()
    for i in range(arr.shape[axis]):
        idx = [slice(None)] * 3
        idx[axis] = i
        sl = arr[tuple(idx)]
        masked = np.isin(sl, values)
        if not masked.any() or masked.all():
            continue
        ys, xs = np.nonzero(~masked)
        vals = sl[~masked]
        new = sl.copy()
        a2 = np.zeros(sl.shape, dtype=bool)
        for (py, px) in zip(*np.nonzero(masked)):
            d2 = (ys - py) ** 2 + (xs - px) ** 2
            m = d2.min()
            cand = np.nonzero(d2 == m)[0]
            new[py, px] = vals[cand[0]]  # np.nonzero order = row-major
            a2[py, px] = len(set(vals[cand].tolist())) > 1
        out[tuple(idx)] = new
        amb[tuple(idx)] = a2
    return out, amb


def brats_label_map(onehot: np.ndarray) -> np.ndarray:
    """ConvertToBratsClassesBasedOnMultiChannel + ChangeLabel3To4 (utils/transforms.py:169-206) on [3, D, H, W]."""
    tc, wt, et = onehot[0].astype(bool), onehot[1].astype(bool), onehot[2].astype(bool)
    lab = np.zeros(tc.shape, dtype=np.uint8)
    lab[et] = 4
    lab[tc & ~et] = 1
    lab[wt & ~tc] = 2
    return lab


def labels_to_channels(label: np.ndarray) -> np.ndarray:
    """MONAI ConvertToMultiChannelBasedOnBratsClasses: (TC, WT, ET)."""
    tc = (label == 1) | (label == 4)
    wt = tc | (label == 2)
    et = label == 4
    return np.stack([tc, wt, et]).astype(np.uint8)


# ------------------------------------------------------------------ seeded inputs shared by the golden generator and tests
def synth_raw(seed, shape=(4, 20, 22, 19)):
    """Raw 'MRI' intensities: positive inside an off-centre ellipsoid, exactly 0 outside, a few interior zeros."""
    g = np.random.default_rng(seed)
    c, d, h, w = shape
    zz, yy, xx = np.meshgrid(np.arange(d), np.arange(h), np.arange(w), indexing="ij")
    inside = ((zz - d * 0.55) / (d * 0.35)) ** 2 + ((yy - h * 0.45) / (h * 0.4)) ** 2 + ((xx - w * 0.5) / (w * 0.3)) ** 2 <= 1
    img = (g.gamma(2.0, 150.0, size=shape) + 5.0).astype(np.float32) * inside[None]
    img[:, d // 2, h // 2, w // 2] = 0.0
    img[min(1, c - 1)] *= (g.random((d, h, w)) > 0.05)
    return img


def synth_labels(seed, shape=(24, 26, 21)):
    """A BraTS-like label map: one big blob with nested labels, a few small islands, a handful of ET voxels."""
    g = np.random.default_rng(seed)
    d, h, w = shape
    zz, yy, xx = np.meshgrid(np.arange(d), np.arange(h), np.arange(w), indexing="ij")
    r = np.sqrt(((zz - d / 2) / (d * 0.3)) ** 2 + ((yy - h / 2) / (h * 0.3)) ** 2 + ((xx - w / 2) / (w * 0.3)) ** 2)
    lab = np.zeros(shape, dtype=np.uint8)
    lab[r < 1.0] = 2
    lab[r < 0.6] = 1
    # islands (sizes 1..14) away from the blob, some touching only diagonally
    lab[1, 1, 1] = 2
    lab[2, 2, 2] = 2  # diagonal neighbour: same component under 26-connectivity
    lab[1:3, h - 4:h - 1, 1:3] = 1  # 12 voxels
    lab[d - 3:d - 1, 1:4, w - 4:w - 2] = 2  # 12 voxels
    lab[d - 2, h - 2, w - 2] = 1
    for _ in range(7):  # rare ET voxels scattered in the core
        p = g.integers(0, 3, size=3) + np.array([d // 2 - 1, h // 2 - 1, w // 2 - 1])
        lab[tuple(p)] = 4
    return lab
