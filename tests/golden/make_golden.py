"""Generates tests/golden/*.npz by running the UNMODIFIED reference (/root/reference, read-only) with the MONAI
shim (oracle/monai_shim.py).  Runs only in the build container; the GPU box consumes the committed fixtures.

    python tests/golden/make_golden.py

Inputs/weights come from oracle/synth.py (name-keyed torch CPU generators), so fixtures hold outputs only.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_loader, synth  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
WIDTH = 16
SHAPE = (32, 32, 32)
VARIANTS = (("instance", "leakyrelu"), ("batch", "elu"), ("none", "relu"), ("group", "leakyrelu"), ("instance", "relu"))


def ramp_predictor(x: torch.Tensor) -> torch.Tensor:
    """A window-position-dependent 'network' so that blending weights matter (3 output channels)."""
    d, h, w = x.shape[2:]
    i = torch.arange(d, dtype=x.dtype).reshape(1, d, 1, 1)
    j = torch.arange(h, dtype=x.dtype).reshape(1, 1, h, 1)
    k = torch.arange(w, dtype=x.dtype).reshape(1, 1, 1, w)
    ramp = 1.0 + 0.01 * (i + 2 * j + 3 * k)
    y = torch.stack([x[:, 0] * 0.5 + x[:, 1], x[:, 2] - x[:, 3], x.sum(1) * 0.25], dim=1)
    return y * ramp.unsqueeze(0)


def main():
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    ref = ref_loader.load()
    x = synth.volume(seed=3, shape=SHAPE)

    # ---- networks
    feats = [WIDTH * 2 ** i for i in range(4)]
    with ref_loader.quiet():
        v1 = ref.equiunet2020.EquiUnet(4, 3, feats, norm_layer="group", act="relu", deep_supervision=True)
        v2 = ref.equiunet2021.EquiUnetASSPEvo(4, 3, feats, norm_layer="group", act="leakyrelu",
                                              deep_supervision=True)
    for ver, net, seed in ((1, v1, 123), (2, v2, 93)):
        params = synth.make_params(ver, WIDTH, seed)
        missing = net.load_state_dict(params, strict=True)
        net.eval()
        with torch.no_grad():
            out, deeps = net(x)
        rec = {"out": out.numpy()}
        for i, dp in enumerate(deeps):
            rec[f"deep{i}_s2"] = dp[..., ::2, ::2, ::2].contiguous().numpy()
        rec["keys"] = np.array(sorted(net.state_dict().keys()))
        np.savez_compressed(os.path.join(HERE, f"net_v{ver}_w{WIDTH}.npz"), **rec)
        print(f"v{ver}: out {tuple(out.shape)} mean {out.mean():.5f} std {out.std():.5f}", missing)

    # ---- the rest of the norm / act factory (networks/factory.py:179-200) on EquiUnet: 16^3 inputs
    rec = {}
    xv = synth.volume(seed=3, shape=(16, 16, 16))
    for norm, act in VARIANTS:
        with ref_loader.quiet():
            net = ref.equiunet2020.EquiUnet(4, 3, feats, norm_layer=norm, act=act, deep_supervision=True)
        net.load_state_dict(synth.make_params(1, WIDTH, 123, norm=norm), strict=True)
        net.eval()
        with torch.no_grad():
            out, deeps = net(xv)
        tag = f"{norm}_{act}"
        rec[f"{tag}_out"] = out.numpy()
        rec[f"{tag}_deep0_s2"] = deeps[0][..., ::2, ::2, ::2].contiguous().numpy()
        rec[f"{tag}_keys"] = np.array(list(net.state_dict().keys()))
        if norm == "batch":  # training-mode forward: batch statistics, running statistics updated
            net.train()
            with torch.no_grad():
                out, _ = net(xv)
            rec[f"{tag}_train_out"] = out.numpy()
            rec[f"{tag}_train_running_mean"] = net.encoder1.ConvBnRelu1.bn.running_mean.numpy().copy()
            rec[f"{tag}_train_running_var"] = net.decoder1.ConvBnRelu2.bn.running_var.numpy().copy()
    np.savez_compressed(os.path.join(HERE, "net_v1_w16_variants.npz"), **rec)

    # ---- sliding window (reference utils/inferers.py), odd sizes, both blend modes, image smaller than roi
    sw = {}
    xs = synth.volume(seed=5, shape=(40, 36, 29))
    for mode in ("constant", "gaussian"):
        for bs in (1, 4):
            y = ref.inferers.sliding_window_inference(xs, (16, 16, 16), bs, ramp_predictor, overlap=0.25, mode=mode)
            sw[f"{mode}_b{bs}"] = y.numpy()
    y = ref.inferers.sliding_window_inference(xs, (48, 32, 16), 2, ramp_predictor, overlap=0.5, mode="gaussian")
    sw["pad_gaussian"] = y.numpy()
    np.savez_compressed(os.path.join(HERE, "sliding_window.npz"), **sw)

    # ---- TTA (reference tta.Compose as configured in src/definer.py:647-658)
    tta = ref.tta
    comp = tta.Compose([tta.OnAxes(axes=["zxy", "xyz"]), tta.HorizontalFlip(), tta.Rotate90(angles=[0, 90, 180, 270])])
    xt = torch.arange(2 * 5 * 5 * 5, dtype=torch.float32).reshape(1, 2, 5, 5, 5)
    rec = {"n": np.array(len(comp))}
    for i, tr in enumerate(comp):
        a = tr.augment_image(xt)
        rec[f"aug{i}"] = a.contiguous().numpy()
        rec[f"rt{i}"] = tr.deaugment_mask(a).contiguous().numpy()
    np.savez_compressed(os.path.join(HERE, "tta16.npz"), **rec)

    # ---- Ranger2020 (reference learning/optimizer.py), 14 steps over the N_sma threshold and two lookaheads
    g = torch.Generator().manual_seed(11)
    p0 = [torch.randn(6, 5, 3, 3, 3, generator=g), torch.randn(7, generator=g), torch.randn(4, 6, generator=g)]
    grads = [[torch.randn(p.shape, generator=g) * 0.1 for p in p0] for _ in range(14)]
    rec = {}
    for tag, use_gc in (("nogc", False), ("gc", True)):
        ps = [torch.nn.Parameter(p.clone()) for p in p0]
        with ref_loader.quiet():
            opt = ref.optimizer.Ranger2020(ps, lr=3e-4, alpha=0.5, k=6, N_sma_threshhold=5, betas=(.95, 0.999),
                                           eps=1e-5, weight_decay=1e-5, use_gc=use_gc, gc_conv_only=False)
        for step in range(14):
            for p, gr in zip(ps, grads[step]):
                p.grad = gr.clone()
            opt.step()
            if step in (4, 5, 13):
                for i, p in enumerate(ps):
                    rec[f"{tag}_s{step + 1}_p{i}"] = p.detach().numpy().copy()
    np.savez_compressed(os.path.join(HERE, "ranger.npz"), **rec)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
