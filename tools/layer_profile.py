"""Per-shape conv timing of one bench step (CUDA events around every conv launch, ops.conv_profile).
python tools/layer_profile.py [v2_tta8|v1_sw|v2_train] -> markdown table on stdout."""
import os
import sys
import warnings
from collections import defaultdict

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from brats21_b200 import engine, networks, ops, synth, tta  # noqa: E402
from brats21_b200.losses import DiceLoss  # noqa: E402
from brats21_b200.optimizer import Ranger2020  # noqa: E402


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "v2_tta8"
    dev = torch.device("cuda:0")
    feats = [48, 96, 192, 384]
    torch.manual_seed(93)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        cls = networks.EquiUnet if wl == "v1_sw" else networks.EquiUnetASSPEvo
        net = cls(4, 3, feats, norm_layer="group", act="relu", deep_supervision=True).to(dev)
    if wl == "v2_train":
        net.train()
        opt = Ranger2020([p for n, p in net.named_parameters() if not n.endswith(".v")], lr=3e-4, weight_decay=1e-5, use_gc=False)
        crit = DiceLoss()
        img = synth.volume(seed=2000, shape=(128, 128, 128)).to(dev)
        tgt = synth.target(shape=(128, 128, 128)).to(dev)
        step = lambda: engine.train_step(None, net, crit, opt, img, tgt)  # noqa: E731
    else:
        net.eval()
        vol = torch.nn.functional.pad(synth.volume(seed=1000, shape=(240, 240, 155)).to(dev), (2, 3))
        comp = tta.get_flip8_transforms() if wl == "v2_tta8" else None
        mode = "gaussian" if wl == "v2_tta8" else "constant"
        step = lambda: engine.predict_volume([net], vol, comp, True, (128, 128, 128), 9 if wl == "v2_tta8" else 4, 0.25, mode)  # noqa: E731
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    step()
    e1.record()
    torch.cuda.synchronize()
    clean_ms = e0.elapsed_time(e1)
    ops.conv_profile = []
    step()
    torch.cuda.synchronize()
    prof, ops.conv_profile = ops.conv_profile, None
    agg = defaultdict(lambda: [0, 0.0, 0.0])
    for a, b, fl, key in prof:
        r = agg[str(key)]
        r[0] += 1
        r[1] += a.elapsed_time(b)
        r[2] += fl
    tot_ms = sum(r[1] for r in agg.values())
    tot_fl = sum(r[2] for r in agg.values())
    print(f"# {wl}: step {clean_ms:.2f} ms; conv launches {len(prof)}, conv time {tot_ms:.2f} ms, "
          f"{tot_fl / tot_ms / 1e9:.1f} TFLOP/s aggregate\n")
    print("| shape key | launches | total ms | share of conv | TFLOP/s |\n|---|---:|---:|---:|---:|")
    for key, r in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {key} | {r[0]} | {r[1]:.2f} | {r[1] / tot_ms:.3f} | {r[2] / r[1] / 1e9:.1f} |")


if __name__ == "__main__":
    main()
