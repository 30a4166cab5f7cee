"""Repeat the first-conv kernel on identical inputs and compare the outputs bit for bit (race detector)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from brats21_b200 import ops  # noqa: E402


def run(shape, cout, reps):
    n, d, h, w = shape
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.zeros((n, d, h, w, 8), device="cuda", dtype=torch.bfloat16)
    x[..., :4] = torch.randn((n, d, h, w, 4), device="cuda", generator=g).to(torch.bfloat16)
    wt = torch.randn((cout, 4, 3, 3, 3), device="cuda", generator=g) / 108 ** 0.5
    b = torch.randn((cout,), device="cuda", generator=g)
    pw = ops.PackedConv(wt, b)
    outs, stats = [], []
    for _ in range(reps):
        st = ops.new_stats(n, "cuda")
        y = torch.full((n, d, h, w, cout), float("nan"), device="cuda", dtype=torch.bfloat16)
        ops.conv3d(x, pw, out=y, stats=st)
        torch.cuda.synchronize()
        outs.append(y)
        stats.append(st.sum(0))
    bad = 0
    for i in range(1, reps):
        diff = (outs[i].float() != outs[0].float()) | torch.isnan(outs[i].float())
        nd = int(diff.sum())
        sd = (stats[i] - stats[0]).abs().max().item() / stats[0].abs().max().item()
        if nd or sd > 1e-6:
            bad += 1
            idx = diff.nonzero()[:6].tolist()
            mx = (outs[i].float() - outs[0].float()).abs().max().item()
            print(f"  rep {i}: {nd} elements differ (max {mx:.3e}), stats rel {sd:.2e}, first {idx}")
    print(f"shape {shape} cout {cout}: {bad} of {reps - 1} repetitions differ from the first", flush=True)


if __name__ == "__main__":
    run((1, 32, 32, 32), 16, 40)
    run((1, 32, 32, 32), 48, 40)
    run((2, 21, 40, 24), 32, 40)
    run((9, 128, 128, 128), 48, 8)
