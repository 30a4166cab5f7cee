"""Generates tests/golden/dice_ce.npz by running the UNMODIFIED reference ``learning.losses.DiceCELoss``
(/root/reference/learning/losses.py:470-595) exactly as ``src/definer.py:204-212`` builds it.  Its cross-entropy half is
reference code; its Dice half calls ``monai.losses.DiceLoss``, which is not vendored and comes from oracle/monai_shim.py
(restated, unpinned — see that file).  Runs only in the build container.

    python tests/golden/make_golden_losses.py
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def inputs():
    g = torch.Generator().manual_seed(23)
    x = torch.randn((2, 3, 6, 5, 7), generator=g) * 2.0
    wt = torch.rand((2, 1, 6, 5, 7), generator=g) > 0.5
    tc = wt & (torch.rand((2, 1, 6, 5, 7), generator=g) > 0.4)
    et = tc & (torch.rand((2, 1, 6, 5, 7), generator=g) > 0.5)
    t = torch.cat([tc, wt, et], dim=1).float()  # MONAI BraTS channel order (TC, WT, ET): nested, multi-label
    return x, t


def main():
    ns = ref_loader.load_engine()
    import learning.losses as losses
    x, t = inputs()
    rec = {}
    for tag, kw in (("default", {}), ("weighted", {"lambda_dice": 0.7, "lambda_ce": 1.3})):
        crit = losses.DiceCELoss(include_background=True, sigmoid=True, softmax=False, squared_pred=True, batch=True,
                                 reduction="mean", **kw)
        xr = x.clone().requires_grad_(True)
        loss = crit(xr, t)
        loss.backward()
        rec[f"{tag}_loss"] = np.array(loss.item(), dtype=np.float64)
        rec[f"{tag}_grad"] = xr.grad.numpy()
        rec[f"{tag}_ce"] = np.array(crit.ce(x, t).item(), dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "dice_ce.npz"), **rec)
    print("wrote dice_ce.npz", {k: float(v) for k, v in rec.items() if v.ndim == 0}, ns.engine.__name__)


if __name__ == "__main__":
    main()
