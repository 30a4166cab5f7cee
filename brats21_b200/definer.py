"""Factories with the reference's names and argument meaning (src/definer.py): ``get_model(args)`` keyed on
``args.model`` with ``args.width / norm / act / dropout / num_classes``; ``get_tta_transforms()``;
``get_activation()``; ``get_post_transforms(args)``; ``make_criterion(args)``; ``make_optimizer(args, model)``.
``get_network`` is an alias of ``get_model`` (the name BASELINE.json uses)."""
from __future__ import annotations

import argparse

import torch

from . import tta as _tta
from .networks import EquiUnet, EquiUnetASSPEvo

_SUPPORTED = {
    "equiunet": EquiUnet,
    "equiunet_assp_evo": EquiUnetASSPEvo,
    "equiunet_assp_evocor": EquiUnetASSPEvo,  # same network, renamed in the reference's docker configs
}


def get_model(args: argparse.Namespace) -> torch.nn.Module:
    """src/definer.py:37-174 for the model families on the accelerated path; NameError otherwise (as the
    reference does for unknown names)."""
    if args.model not in _SUPPORTED:
        raise NameError("Not Supported Model")
    kwargs = {"inplanes": 4,
              "num_classes": args.num_classes,
              "features": [args.width * 2 ** i for i in range(4)],
              "norm_layer": getattr(args, "norm", "group"),
              "act": getattr(args, "act", "relu"),
              "deep_supervision": True,  # hard-coded in the reference (definer.py:140)
              "dropout": getattr(args, "dropout", 0)}
    return _SUPPORTED[args.model](**kwargs)


get_network = get_model


def get_tta_transforms() -> _tta.Compose:
    return _tta.get_tta_transforms()


def get_activation():
    """monai Activations(sigmoid=True) (definer.py:661-668)."""
    return torch.sigmoid


def get_post_transforms(args: argparse.Namespace):
    """src/definer.py:671-697.  Plain case: AsDiscrete(threshold_values=True, logit_thresh).  With
    ``args.cleaning_areas`` / ``args.replace_value``: threshold -> BraTS label map (3 -> 4) -> connected-component
    size filter -> rare-label replacement -> back to (TC, WT, ET) channels, all on the GPU (postprocess.py)."""
    thresh = getattr(args, "logit_threshold", 0.5)
    cleaning = getattr(args, "cleaning_areas", False)
    replace = getattr(args, "replace_value", False)
    if not (cleaning or replace):
        return lambda x: (x >= thresh).float()
    from . import ops, postprocess

    def post(prob):
        if prob.dim() != 5 or prob.shape[0] != 1 or prob.shape[1] != 3:
            raise AssertionError("Number of channel need to be 3 (TC/WT/ET), batch 1")
        _, label = ops.labels_finalize(prob[0].contiguous().float(), 1.0, thresh, want_onehot=False)
        label = label[None, None]
        if cleaning:
            label = postprocess.KeepLargestConnectedComponent(getattr(args, "cleaning_areas_threshold", 10))(label)
        if replace:
            label = postprocess.ReplaceWithClosestValue(labels=[3], thresh=getattr(args, "replace_value_threshold", 20))(label)
        return postprocess.labels_to_channels(label).float()

    return post


def make_criterion(args: argparse.Namespace):
    from .losses import DiceLoss
    if args.criterion == "dice":
        return DiceLoss(jaccard=False)
    if args.criterion == "jaccard":
        return DiceLoss(jaccard=True)
    if args.criterion == "dice_ce":  # src/definer.py:204-212
        from .losses import DiceCELoss
        return DiceCELoss(include_background=True, sigmoid=True, softmax=False, squared_pred=True, batch=True,
                          reduction="mean")
    raise NameError("Not Supported Criterion")


def make_optimizer(args: argparse.Namespace, model: torch.nn.Module):
    from .optimizer import Ranger2020
    if args.optimizer != "ranger":
        raise NameError("Not Supported Optimizer")
    trainable = [p for p in model.parameters() if p.requires_grad]
    return Ranger2020(trainable, lr=args.learning_rate, alpha=0.5, k=6, N_sma_threshhold=5, betas=(.95, 0.999),
                      eps=1e-5, weight_decay=args.weight_decay, use_gc=getattr(args, "use_gc", False),
                      use_gcnorm=getattr(args, "use_gcnorm", False), normloss=getattr(args, "normloss", False),
                      normloss_factor=getattr(args, "normloss_factor", 1e-4),
                      gc_conv_only=getattr(args, "gc_conv_only", False), gc_loc=True)
