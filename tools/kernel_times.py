"""Per-kernel GPU time of one bench step from CUPTI (torch.profiler): warm, in-order, no replay — complements the
ncu launch list (cold-cache, serialised).  python tools/kernel_times.py [v2_tta8|v1_sw|v2_train] -> markdown."""
import os
import sys
import warnings
from collections import defaultdict

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from brats21_b200 import engine, networks, synth, tta  # noqa: E402
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
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    step()
    e1.record()
    torch.cuda.synchronize()
    step_ms = e0.elapsed_time(e1)
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        step()
        torch.cuda.synchronize()
    agg = defaultdict(lambda: [0, 0.0])
    for ev in prof.events():
        if ev.device_type is not None and "cuda" in str(ev.device_type).lower() and ev.device_time_total > 0:
            name = ev.name.split("(")[0].replace("void ", "")
            agg[name][0] += 1
            agg[name][1] += ev.device_time_total / 1e3
    tot = sum(v[1] for v in agg.values())
    print(f"# {wl}: step {step_ms:.2f} ms (CUDA events, unprofiled); kernels under CUPTI: {tot:.2f} ms over "
          f"{sum(v[0] for v in agg.values())} launches\n")
    print("| kernel | launches | total ms | share | avg us |\n|---|---:|---:|---:|---:|")
    for name, (cnt, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        print(f"| `{name[:90]}` | {cnt} | {ms:.2f} | {ms / tot:.3f} | {1e3 * ms / cnt:.1f} |")


if __name__ == "__main__":
    main()
