"""Repeat the TRAINING forward of a w16 network on identical inputs and report the first tape tensor that differs
from the first run (bitwise): locates run-to-run non-determinism of the forward pass."""
import os
import sys
import warnings

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from brats21_b200 import networks, ops  # noqa: E402
from oracle import synth  # noqa: E402

DEV = torch.device("cuda:0")
ORDER = ["encoder1", "encoder2", "encoder3", "encoder4", "aspp.conv_k1", "bridge1", "bridge2", "bridge3", "upconv3",
         "decoder3", "upconv2", "decoder2", "upconv1", "decoder1"]


def snapshot(net, x):
    with torch.no_grad():
        x8 = net.pack_input(x)
        out, deeps, tape = net._forward_train(x8, True)
    torch.cuda.synchronize()
    snap = [("x8", x8.clone())]
    for name in ORDER:
        t = tape[name]
        for k in ("z0", "st0", "a0", "z1", "st1", "y", "z", "st"):
            if k in t:
                v = t[k]
                snap.append((f"{name}.{k}", (v.sum(0) if v.dtype == torch.float64 else v).clone()))
    snap.append(("out", out.clone()))
    return snap


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    width = 16
    params = {k: v.to(DEV) for k, v in synth.make_params(2, width, 93).items()}
    x = synth.volume(seed=10, shape=(32, 32, 32)).to(DEV)
    for use in (True, False):
        ops.use_input = use
        for fresh in (False, True):
            net = None
            ref = None
            bad = 0
            for r in range(reps):
                if net is None or fresh:
                    with warnings.catch_warnings():
                        warnings.simplefilter("ignore")
                        net = networks.EquiUnetASSPEvo(4, 3, [width * 2 ** i for i in range(4)], deep_supervision=True).to(DEV)
                    net.load_state_dict(params)
                    net.train()
                snap = snapshot(net, x)
                if ref is None:
                    ref = snap
                    continue
                for (name, a), (_, b) in zip(snap, ref):
                    same = torch.equal(a, b) if a.dtype != torch.float64 else bool(((a - b).abs() <= 1e-6 * b.abs().max()).all())
                    if not same:
                        d = (a.double() - b.double()).abs()
                        print(f"  use_input={use} fresh_net={fresh} rep {r}: first difference at {name}: "
                              f"{int((d > 0).sum())} elements, max {d.max().item():.3e}", flush=True)
                        bad += 1
                        break
            print(f"use_input={use} fresh_net={fresh}: {bad} of {reps - 1} forwards differ", flush=True)


if __name__ == "__main__":
    main()
