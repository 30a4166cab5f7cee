"""Generates tests/golden/prepost.npz by running the UNMODIFIED reference functions of utils/transforms.py
(/root/reference, read-only) on small seeded inputs.  Runs only in the build container.

    python tests/golden/make_golden_prepost.py

The module's third-party imports that are absent from this image are stubbed for IMPORT ONLY: SimpleITK and
monai.{config,transforms} are never called by the functions exercised here; ``skimage.morphology.label`` IS called by
get_largest_component and is provided by scipy.ndimage.label with the full-connectivity structuring element
(skimage's default connectivity = ndim), so that part of the fixture is "reference code over a restated label()".
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402
from oracle.prepost import synth_labels, synth_raw  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def _stub_modules():
    from scipy import ndimage
    ref_loader.load()  # MONAI shim + collections aliases + sys.path
    if not hasattr(np, "int"):
        np.int = int  # utils/transforms.py:503 uses the alias removed in numpy 1.24

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    mod("SimpleITK")
    cfg = mod("monai.config", DtypeLike=object, KeysCollection=object)
    sys.modules["monai"].config = cfg

    class Transform:
        pass

    class MapTransform(Transform):
        def __init__(self, keys, *a, **k):
            self.keys = keys

    tr = sys.modules["monai.transforms"]
    tr.Transform, tr.MapTransform, tr.BorderPad = Transform, MapTransform, object

    def label(mask, *a, **k):
        return ndimage.label(mask, structure=np.ones((3,) * mask.ndim, dtype=bool))[0]

    sk = mod("skimage")
    sk.morphology = mod("skimage.morphology", label=label)


def main():
    _stub_modules()
    import utils.transforms as T  # the unmodified reference module
    rec = {}
    for i, seed in enumerate((0, 1)):
        img = synth_raw(seed)
        for ro in (False, True):
            norm = T.NormalizeIntensity(nonzero=True, channel_wise=True, remove_outliers=ro)
            out = norm(img.copy())
            padded, pb, pa = T.shape_to_divisible(torch.from_numpy(out)[None], k=8)
            rec[f"norm{i}_ro{int(ro)}"] = padded[0].numpy()
            rec[f"norm{i}_pb"] = np.asarray(pb)
            rec[f"norm{i}_pa"] = np.asarray(pa)
    for i, seed in enumerate((3, 4)):
        lab = synth_labels(seed)
        for thr in (None, 1, 10, 12, 100000):
            out = T.get_largest_component(lab.copy()[None, None], threshold=thr)
            rec[f"cc{i}_t{thr}"] = np.asarray(out)[0, 0]
        for axis in (0, 1, 2):
            tr = T.ReplaceWithClosestValue(labels=[3], thresh=20, axis=axis)
            out = tr(torch.from_numpy(lab.copy())[None, None])
            rec[f"rep{i}_a{axis}"] = np.asarray(out)[0, 0].astype(np.uint8)
        onehot = np.stack([(lab == 1) | (lab == 4), lab > 0, lab == 4]).astype(np.float32)
        conv = T.ConvertToBratsClassesBasedOnMultiChannel()(torch.from_numpy(onehot)[None])
        rec[f"brats{i}"] = T.ChangeLabel3To4()(conv)[0, 0].numpy().astype(np.uint8)
    np.savez_compressed(os.path.join(HERE, "prepost.npz"), **rec)
    print("wrote", os.path.join(HERE, "prepost.npz"), sorted(rec))


if __name__ == "__main__":
    main()
