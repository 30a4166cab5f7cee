"""Step bodies of the reference ``Engine`` (learning/engine.py) on the B200 path.

Drop-in pieces (same names, argument meaning and return structure as the reference):
  * ``compute_output(args, model, img, is_train)``   Engine._compute_output   (engine.py:298-310)
  * ``compute_loss(args, criterion, outputs, label)`` Engine._compute_loss     (engine.py:312-333)
  * ``apply_tta(args, model, img, tta_transforms)``   Engine._apply_tta        (engine.py:424-440)

Fused fast path (what bench.py measures): ``predict_volume`` runs models x TTA variants x windows entirely on the
GPU — windows are read out of the source volume through the variant's signed permutation, logits are blended
with fp32 accumulators, de-augmented, sigmoid-ed and summed in place, and one final pass thresholds the mean and
emits the BraTS label map with background removal (engine.py:236-252, utils/transforms.py:169-206,536-550).
Nothing is copied to the host until the final uint8 labels.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch

from . import ops
from .inferers import WindowPlan, accumulate_windows, sliding_window_inference
from .tta import Compose, get_tta_transforms


def _flatten(x):
    if isinstance(x, (list, tuple)):
        out = []
        for y in x:
            out.extend(_flatten(y))
        return out
    return [x]


def _apply_f(seq, f):
    if isinstance(seq, (list, tuple)):
        return [_apply_f(s, f) for s in seq]
    return f(seq)


def compute_output(args, model, img, is_train: bool = False):
    """autocast is a no-op here: the kernels already compute in bf16 with fp32 accumulation."""
    if getattr(args, "sliding_window_inference", False) and not is_train:
        return sliding_window_inference(img, roi_size=args.sliding_window_size, sw_batch_size=1, predictor=model,
                                        device=torch.device("cpu"))
    return model(img)


def compute_loss(args, criterion, outputs, label=None):
    if isinstance(outputs, (tuple, list)):  # deep supervision
        heads = _flatten(outputs)
        loss = torch.mean(torch.stack([criterion(h, label) for h in heads])) if label is not None else None
        return heads[0], loss
    return outputs, (criterion(outputs, label) if label is not None else None)


def train_step(args, model, criterion, optimizer, image, target, scaler=None):
    """Body of the reference training loop for one batch (learning/engine.py:104-122): zero_grad -> forward ->
    deep-supervision loss -> backward -> optimizer step.  Returns the loss as a DEVICE tensor: the reference's
    per-step ``loss.item()`` (engine.py:114) is a host sync the caller can do when it wants the number."""
    model.zero_grad()
    outputs = compute_output(args, model, image, is_train=True)
    _, loss = compute_loss(args, criterion, outputs, target)
    if scaler is not None:  # torch.cuda.amp.GradScaler as in main_train.py:110 (not needed with bf16)
        scaler.scale(loss).backward()
        scaler.step(optimizer)
        scaler.update()
    else:
        loss.backward()
        optimizer.step()
    return loss.detach()


class TrainStep:
    """``train_step`` replayed from ONE CUDA graph per input shape: zero_grad, bf16 weight re-pack, forward, Dice over
    all heads, the hand-scheduled backward (weight gradients on their side stream, the data-parallel bucket
    all-reduces on theirs), and the fused Ranger launch — ~450 launches, several memsets and the autograd bookkeeping
    become one ``cudaGraphLaunch``.  The step-dependent optimizer scalars (rectified step size, look-ahead phase,
    learning rate) are uploaded to a small device buffer before every replay (``Ranger2020.begin_graph_step``), so
    schedulers keep working; image / target are copied into static buffers.

        step = TrainStep(model, criterion, optimizer)
        for batch in loader:
            loss = step(batch["img"].cuda(non_blocking=True), batch["seg"].cuda(non_blocking=True))   # device scalar

    The optimizer must be ``brats21_b200.optimizer.Ranger2020``.  Same arithmetic as ``train_step`` (same kernels in
    the same order); ``tests/test_gpu_train.py::test_graphed_train_step_matches_eager`` pins that."""

    def __init__(self, model, criterion, optimizer, args=None, eager_warmup: int = 2):
        if not hasattr(optimizer, "begin_graph_step"):
            raise TypeError("TrainStep needs brats21_b200.optimizer.Ranger2020 (device-resident step scalars)")
        self.model, self.criterion, self.optimizer, self.args = model, criterion, optimizer, args
        self.eager_warmup = max(int(eager_warmup), 1)
        self._graphs = {}
        self.eager_steps = 0

    def _params(self):
        return [p for g in self.optimizer.param_groups for p in g["params"]]

    def _capture(self, image, target):
        dev = image.device
        img = torch.empty_like(image, dtype=torch.float32).contiguous()
        tgt = torch.empty_like(target, dtype=torch.float32).contiguous()
        torch.cuda.synchronize(dev)
        self.optimizer.enable_graph_mode()
        graph = torch.cuda.CUDAGraph()
        l0 = ops._lib.launch_count
        try:
            with torch.cuda.graph(graph):  # records only: nothing executes, no host state advances (graph mode)
                loss = train_step(self.args, self.model, self.criterion, self.optimizer, img, tgt)
        except Exception:
            self.optimizer.disable_graph_mode()
            raise
        return (graph, img, tgt, loss, ops._lib.launch_count - l0)

    def __call__(self, image: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        if not image.is_cuda:
            raise RuntimeError("TrainStep runs on CUDA only (no CPU fallback)")
        key = (tuple(image.shape), tuple(target.shape), image.device.index)
        entry = self._graphs.get(key)
        if entry is None or isinstance(entry, int):
            done = entry or 0
            if done < self.eager_warmup:
                # the first steps of a shape run eagerly: optimizer state, workspaces, packed weights, NCCL
                # communicators and kernel attributes must exist before a capture
                self.optimizer.disable_graph_mode()
                self._graphs[key] = done + 1
                self.eager_steps += 1
                return train_step(self.args, self.model, self.criterion, self.optimizer, image, target)
            entry = self._graphs[key] = self._capture(image, target)
        graph, img, tgt, loss, launches = entry
        self.optimizer.enable_graph_mode()
        img.copy_(image, non_blocking=True)
        tgt.copy_(target, non_blocking=True)
        self.optimizer.begin_graph_step()
        graph.replay()
        ops._lib.launch_count += launches
        # the replayed optimizer launch wrote the parameters: packed-weight caches outside the graph key on _version
        torch._C._increment_version(self._params())
        return loss


def apply_tta(args, model, img, tta_transforms: Optional[Compose]) -> List:
    """Drop-in Engine._apply_tta: list (one entry per variant) of de-augmented outputs moved to the CPU."""
    outs = []
    for tr in tta_transforms:
        o = compute_output(args, model, tr.augment_image(img))
        outs.append(_apply_f(o, lambda x: tr.deaugment_mask(x).cpu()))
    return outs


@torch.no_grad()
def predict_volume(models: Sequence, image: torch.Tensor, tta_transforms: Optional[Compose] = None,
                   sliding_window: bool = True, roi_size=(128, 128, 128), sw_batch_size: int = 4,
                   overlap: float = 0.25, mode: str = "gaussian", sigma_scale: float = 0.125,
                   logit_thresh: float = 0.5, remove_background: bool = True, return_prob: bool = False):
    """Ensemble x TTA x sliding-window inference of ONE volume, fused on the GPU.

    image: [1, C, D, H, W] fp32 on CUDA, spatial dims already divisible by 8 (shape_to_divisible).
    Returns (onehot uint8 [1, 3, D, H, W] in (TC, WT, ET) order, label uint8 [1, 1, D, H, W]) and, optionally, the
    mean probability map.  Equivalent reference flow: Engine.evaluate(use_tta=...) -> post_trans -> labels.
    """
    if image.dim() != 5 or image.shape[0] != 1:
        raise ValueError("predict_volume takes one volume: [1, C, D, H, W]")
    vol = image.detach().to(torch.float32).contiguous()
    vdims = tuple(vol.shape[2:])
    if tta_transforms is None:
        variants = [((0, 1, 2), (0, 0, 0))]
    else:
        variants = []
        for tr in tta_transforms:
            if tr.variant is None:
                raise ValueError("fused TTA needs signed-permutation transforms; use apply_tta for others")
            variants.append(tr.variant)
    k = models[0].num_classes
    prob_sum = torch.zeros((k,) + vdims, dtype=torch.float32, device=vol.device)
    acc_cache = {}
    count = 0
    for model in models:
        model._ensure_packed()
        for perm, flip in variants:
            adims = [0, 0, 0]
            for j in range(3):
                adims[perm[j]] = vdims[j]
            if sliding_window:
                plan = WindowPlan(adims, roi_size, overlap, mode, sigma_scale, vol.device)
                key = plan.image_size
                if key not in acc_cache:
                    acc_cache[key] = torch.empty((k,) + plan.image_size, dtype=torch.float32, device=vol.device)
                acc = acc_cache[key]
                acc.zero_()
                accumulate_windows(vol, 0, model, plan, sw_batch_size, acc, (perm, flip))
                ops.tta_accumulate(acc, plan.count, prob_sum, perm, flip, pad_before=plan.pad_before)
            else:
                # full-volume forward (the reference's default path, engine.py:309)
                ws = model._ws.setdefault(("full",) + tuple(adims), {})
                x8 = model._buf(ws, "x8", (1,) + tuple(adims) + (8,))
                ops.pack_windows(vol, x8, [(0, 0, 0)], perm=perm, flip=flip)
                logits = model.forward_infer(x8)
                ops.tta_accumulate(logits[0], None, prob_sum, perm, flip)
            count += 1
    onehot, label = ops.labels_finalize(prob_sum, count, logit_thresh, image=vol[0] if remove_background else None)
    if return_prob:
        return onehot[None], label[None, None], prob_sum / count
    return onehot[None], label[None, None]


@torch.no_grad()
def predict_case(models: Sequence, raw_image: torch.Tensor, tta_transforms: Optional[Compose] = None,
                 roi_size=(128, 128, 128), sw_batch_size: int = 4, overlap: float = 0.25, mode: str = "gaussian",
                 logit_thresh: float = 0.5, cleaning_areas_threshold: Optional[int] = None,
                 replace_value_threshold: Optional[int] = None, remove_outliers: bool = False):
    """Raw intensities -> BraTS label map at the ORIGINAL image size, everything on the GPU: the reference's
    inference flow src/definer.py:561-567 (CropForegroundd, NormalizeIntensityd) -> learning/engine.py:226-252
    (shape_to_divisible, ensemble x TTA x sliding window, mean, threshold, background removal) -> post transforms
    (definer.py:679-692, optional) -> shape_to_original + pad_back_to_shape_before_compose (engine.py:258-262).

    raw_image: [C, D, H, W] fp32 on CUDA.  Returns uint8 [D, H, W] with labels 0/1/2/4.
    """
    from . import postprocess, preprocess
    vol, meta = preprocess.crop_normalize_pad(raw_image, 8, remove_outliers=remove_outliers)
    post = cleaning_areas_threshold is not None or replace_value_threshold is not None
    # reference order (learning/engine.py:249-256): threshold -> post transforms -> remove_background_voxels; without
    # post transforms the background mask is fused into the label pass
    _, label = predict_volume(models, vol, tta_transforms, True, roi_size, sw_batch_size, overlap, mode,
                              logit_thresh=logit_thresh, remove_background=not post)
    if cleaning_areas_threshold is not None:
        label = postprocess.KeepLargestConnectedComponent(cleaning_areas_threshold)(label)
    if replace_value_threshold is not None:
        label = postprocess.ReplaceWithClosestValue(labels=[3], thresh=replace_value_threshold)(label)
    if post:
        label = ops.mask_background(label.contiguous(), vol[0])
    return postprocess.pad_back(label[0, 0], meta)
