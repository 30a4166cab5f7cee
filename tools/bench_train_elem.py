"""Time the HBM-bound kernels of the training step on single tensors (CUDA events, L2 flushed between runs):
    python tools/bench_train_elem.py [c,n,d,h,w ...]        default: the level-1/2/3 shapes of the width-48 V2 step
Per kernel: ms, algorithmic GB/s (bf16 activations read / written once) and the fraction of the measured HBM peak."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from brats21_b200 import ops  # noqa: E402
from tools.bench_conv import time_it  # noqa: E402

PEAK = 6535.4
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:  # noqa: BLE001
    pass


def report(name, shape, ms, nbytes):
    gbs = nbytes / ms / 1e6
    print(f"| {name} | {shape} | {ms * 1e3:.1f} | {nbytes / 1e6:.1f} | {gbs:.0f} | {gbs / PEAK:.2f} |", flush=True)


def main():
    specs = sys.argv[1:] or ["48,1,128,128,128", "24,1,128,128,128", "96,1,64,64,64", "24,1,64,64,64", "192,1,32,32,32",
                             "384,1,16,16,16"]
    print("| kernel | c,n,d,h,w | us | algorithmic MB | GB/s | of HBM peak |\n|---|---|---:|---:|---:|---:|")
    for spec in specs:
        c, n, d, h, w = [int(v) for v in spec.split(",")]
        g = torch.Generator(device="cuda").manual_seed(3)
        z = torch.randn((n, d, h, w, c), device="cuda", generator=g).to(torch.bfloat16)
        dy = torch.randn((n, d, h, w, c), device="cuda", generator=g).to(torch.bfloat16)
        y = torch.empty_like(z)
        dz = torch.empty_like(z)
        nel = z.numel()
        stats = torch.zeros((ops._lib.STAT_SLOTS, n, 8, 2), dtype=torch.float64, device="cuda")
        zf = z.float().reshape(n, -1, 8, c // 8)
        stats[0, :, :, 0] = zf.sum(dim=(1, 3)).double()
        stats[0, :, :, 1] = (zf * zf).sum(dim=(1, 3)).double()
        gamma = torch.ones(c, device="cuda")
        beta = torch.zeros(c, device="cuda")
        dgamma, dbeta, colsum = torch.zeros(c, device="cuda"), torch.zeros(c, device="cuda"), torch.zeros(c, device="cuda")
        csum = torch.zeros((n, c), device="cuda")
        ws = ops.norm_bwd_workspace(n, c, "cuda")
        hid = c // 2
        se = dict(scale=torch.full((n, c), 1.5, device="cuda"), mean=torch.zeros((n, c), device="cuda"),
                  w1=torch.randn((hid, c), device="cuda") * 0.1, b1=torch.zeros(hid, device="cuda"),
                  w2=torch.randn((c, hid), device="cuda") * 0.1, b2=torch.zeros(c, device="cuda"),
                  dw1=torch.zeros((hid, c), device="cuda"), db1=torch.zeros(hid, device="cuda"),
                  dw2=torch.zeros((c, hid), device="cuda"), db2=torch.zeros(c, device="cuda"))
        for mode, tag in ((ops.EVO_S0, "evo"), (ops.GN_RELU, "gn")):
            ms = time_it(lambda: ops.norm_apply(z, stats, gamma, beta, mode, out=y))
            report(f"norm_apply {tag}", spec, ms, 4.0 * nel)
            ms = time_it(lambda: ops.norm_apply(z, stats, gamma, beta, mode, out=y, chan_sum=csum))
            report(f"norm_apply {tag} + channel sums", spec, ms, 4.0 * nel)
            ms = time_it(lambda: ops.norm_bwd(dy, z, dz, stats, gamma, beta, dgamma, dbeta, mode, colsum=colsum,
                                              workspace=ws))
            report(f"norm_bwd {tag} (reduce + coeffs + apply)", spec, ms, 10.0 * nel)
            if mode == ops.EVO_S0:
                ms = time_it(lambda: ops.norm_bwd(dy, z, dz, stats, gamma, beta, dgamma, dbeta, mode, colsum=colsum,
                                                  se=se, workspace=ws))
                report(f"norm_bwd {tag} + SE", spec, ms, 10.0 * nel)
        if d % 2 == 0:
            pooled = torch.empty((n, d // 2, h // 2, w // 2, 2 * c), device="cuda", dtype=torch.bfloat16)
            dp = torch.randn_like(pooled)
            ms = time_it(lambda: ops.pool_bwd(z, dp, dz, 2, add=dy))
            report("pool_bwd (max|avg, + add)", spec, ms, 6.0 * nel + 2.0 * dp.numel())
            dx = torch.empty((n, d // 2, h // 2, w // 2, c), device="cuda", dtype=torch.bfloat16)
            ms = time_it(lambda: ops.upsample2x_bwd(dy, dx))
            report("upsample2x_bwd", spec, ms, 2.0 * nel + 2.0 * dx.numel())
            ms = time_it(lambda: ops.upsample2x(dx, y))
            report("upsample2x", spec, ms, 2.0 * nel + 2.0 * dx.numel())
        if c <= 192:
            wt = torch.randn((3, c), device="cuda")
            dl = torch.randn((n, 3, d, h, w), device="cuda")
            ms = time_it(lambda: ops.head_conv_bwd(z, wt, dl, dz))
            report("head_conv_bwd", spec, ms, 4.0 * nel + 12.0 * n * d * h * w)
        ms = time_it(lambda: ops.add_inplace(dz, dy))
        report("add_inplace", spec, ms, 6.0 * nel)


if __name__ == "__main__":
    main()
