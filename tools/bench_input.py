"""Time the first-conv kernel (csrc/conv_input.cu) against the plane-marching kernel on one window batch
(CUDA events, L2 flushed):  python tools/bench_input.py [n d h w]   (default 9 128 128 128)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from brats21_b200 import ops  # noqa: E402
from tools.bench_conv import time_it  # noqa: E402


def main():
    n, d, h, w = [int(v) for v in sys.argv[1:5]] if len(sys.argv) >= 5 else (9, 128, 128, 128)
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.zeros((n, d, h, w, 8), device="cuda", dtype=torch.bfloat16)
    x[..., :4] = torch.randn((n, d, h, w, 4), device="cuda", generator=g).to(torch.bfloat16)
    wt = torch.randn((48, 4, 3, 3, 3), device="cuda", generator=g) / 108 ** 0.5
    b = torch.randn((48,), device="cuda", generator=g)
    pw = ops.PackedConv(wt, b)
    st = ops.new_stats(n, "cuda")
    y = torch.empty((n, d, h, w, 48), device="cuda", dtype=torch.bfloat16)
    for use in (True, False):
        ops.use_input = use
        for act in (False, True):
            ms = time_it(lambda: ops.conv3d_fold(x, pw, y, st, ab=None, act=act))
            gb = 2.0 * n * d * h * w * (4 + 48) / 1e9
            print(f"{'input' if use else 'march'} act={int(act)} variant={os.environ.get('B21_INPUT_VARIANT', '0')}: "
                  f"{ms:.4f} ms  {gb / ms * 1e3:.0f} GB/s algorithmic", flush=True)


if __name__ == "__main__":
    main()
