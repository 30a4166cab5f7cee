"""Time single conv launches on the B200 (CUDA events, L2 flushed between launches by the >126 MB working set or an
explicit flush).  python tools/bench_conv.py cin,cout,k,dil,n,d,h,w[,flags] ...   flags: notap nomarch nopoint noslide"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from brats21_b200 import ops  # noqa: E402


def time_it(fn, reps=5):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def main():
    for spec in sys.argv[1:]:
        parts = spec.split(",")
        cin, cout, k, dil, n, d, h, w = [int(v) for v in parts[:8]]
        flags = parts[8:]
        g = torch.Generator(device="cuda").manual_seed(1)
        x = torch.randn((n, d, h, w, cin), device="cuda", generator=g).to(torch.bfloat16)
        wt = torch.randn((cout, cin, k, k, k), device="cuda", generator=g) / (cin * k ** 3) ** 0.5
        b = torch.randn((cout,), device="cuda", generator=g)
        pw = ops.PackedConv(wt, b)
        st = ops.new_stats(n, "cuda")
        y = torch.empty((n, d, h, w, cout), device="cuda", dtype=torch.bfloat16)
        for name in ("march", "point", "slide"):
            if hasattr(ops, "use_" + name):
                setattr(ops, "use_" + name, ("no" + name) not in flags)
        ms = time_it(lambda: ops.conv3d(x, pw, out=y, stats=st, dil=dil))
        flops = 2.0 * n * d * h * w * cin * cout * k ** 3
        bytes_ = 2.0 * n * d * h * w * (cin + cout)
        print(f"{spec}: {ms:.4f} ms  {flops / ms / 1e9:.1f} TFLOP/s  {bytes_ / ms / 1e6:.1f} GB/s (algorithmic r+w)",
              flush=True)


if __name__ == "__main__":
    main()
