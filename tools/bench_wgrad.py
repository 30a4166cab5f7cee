"""Time single wgrad launches: python tools/bench_wgrad.py cin,cout,n,d,h,w ..."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from brats21_b200 import ops  # noqa: E402
from tools.bench_conv import time_it  # noqa: E402

for spec in sys.argv[1:]:
    parts = spec.split(",")
    cin, cout, n, d, h, w = [int(v) for v in parts[:6]]
    ops.use_wgrad_march = "nomarch" not in parts[6:]
    x = torch.randn((n, d, h, w, cin), device="cuda").to(torch.bfloat16)
    dz = torch.randn((n, d, h, w, cout), device="cuda").to(torch.bfloat16)
    dw = torch.zeros((cout, cin, 3, 3, 3), device="cuda")
    ms = time_it(lambda: ops.conv3d_wgrad(x, dz, dw))
    fl = 2.0 * n * d * h * w * cin * cout * 27
    print(f"wgrad {spec}: {ms:.4f} ms  {fl / ms / 1e9:.1f} TFLOP/s", flush=True)
