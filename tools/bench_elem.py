"""Time the HBM-bound kernels at the sizes of one EquiUNet-ASPP-Evo window batch (4 x 128^3) on the B200.
python tools/bench_elem.py  -> name, ms, achieved GB/s (algorithmic read+write bytes)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from brats21_b200 import ops  # noqa: E402

DEV = "cuda"


def time_it(fn, reps=5):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
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


def report(name, ms, nbytes):
    print(f"{name}: {ms:.4f} ms  {nbytes / ms / 1e6:.0f} GB/s", flush=True)


def main():
    n = 4
    bf = torch.bfloat16
    # upsample 24 ch 64^3 -> 128^3 into the upper half of a 48-channel concat buffer
    x = torch.randn((n, 64, 64, 64, 24), device=DEV).to(bf)
    cat = torch.zeros((n, 128, 128, 128, 48), device=DEV, dtype=bf)
    report("upsample2x 24ch 64^3->128^3 (b4)", time_it(lambda: ops.upsample2x(x, cat[..., 24:])),
           x.numel() * 2 + cat.numel())
    dense = torch.zeros((n, 128, 128, 128, 24), device=DEV, dtype=bf)
    report("upsample2x 24ch 64^3->128^3 (b4) into a DENSE 24-ch buffer", time_it(lambda: ops.upsample2x(x, dense)),
           x.numel() * 2 + dense.numel() * 2)
    x32 = torch.randn((n, 64, 64, 64, 32), device=DEV).to(bf)
    cat64 = torch.zeros((n, 128, 128, 128, 64), device=DEV, dtype=bf)
    report("upsample2x 32ch 64^3->128^3 (b4) into half of a 64-ch buffer (sector aligned)",
           time_it(lambda: ops.upsample2x(x32, cat64[..., 32:])), x32.numel() * 2 + cat64.numel())
    x2 = torch.randn((n, 32, 32, 32, 48), device=DEV).to(bf)
    cat2 = torch.zeros((n, 64, 64, 64, 96), device=DEV, dtype=bf)
    report("upsample2x 48ch 32^3->64^3 (b4)", time_it(lambda: ops.upsample2x(x2, cat2[..., 48:])),
           x2.numel() * 2 + cat2.numel())
    # affine max/avg pool 48 ch 128^3 -> 96 ch 64^3
    s = torch.randn((n, 128, 128, 128, 48), device=DEV).to(bf)
    ab = torch.randn((2, n, 48), device=DEV)
    pooled = torch.empty((n, 64, 64, 64, 96), device=DEV, dtype=bf)
    report("affine_pool 48ch 128^3 (b4)", time_it(lambda: ops.affine_pool(s, ab[0], ab[1], pooled, 2)),
           s.numel() * 2 + pooled.numel() * 2)
    # norm_apply EvoNorm 48 ch 128^3 in place (explicit path / training)
    st = torch.zeros((32, n, 8, 2), dtype=torch.float64, device=DEV)
    st[0, :, :, 0] = 0.0
    st[0, :, :, 1] = float(128 ** 3 * 6)
    g, b = torch.ones(48, device=DEV), torch.zeros(48, device=DEV)
    report("norm_apply evo 48ch 128^3 (b4)", time_it(lambda: ops.norm_apply(s, st, g, b, ops.EVO_S0)), s.numel() * 4)
    report("norm_apply gn+relu 48ch 128^3 (b4)", time_it(lambda: ops.norm_apply(s, st, g, b, ops.GN_RELU)), s.numel() * 4)
    # head conv 48 -> 3
    wh, bh = torch.randn((3, 48), device=DEV), torch.randn(3, device=DEV)
    out = torch.empty((n, 3, 128, 128, 128), device=DEV)
    report("head_conv 48->3 128^3 (b4)", time_it(lambda: ops.head_conv(s, wh, bh, scale=ab[0], offset=ab[1], out=out)),
           s.numel() * 2 + out.numel() * 4)
    # tiny per-conv launches of the folded path
    wt = torch.randn((96, 96, 3, 3, 3), device=DEV) / 50
    pw = ops.PackedConv(wt, torch.zeros(96, device=DEV))
    ab96 = torch.randn((2, n, 96), device=DEV)
    ops._fold_prepare(pw, ab96[0], ab96[1])
    report("fold_prepare 96->96 (pack x4 + bias table)", time_it(lambda: ops._fold_prepare(pw, ab96[0], ab96[1])), 1)
    st96 = torch.ones((32, n, 8, 2), dtype=torch.float64, device=DEV)
    cs = torch.randn((n, 96), device=DEV)
    se = (torch.randn((48, 96), device=DEV), torch.randn(48, device=DEV), torch.randn((96, 48), device=DEV),
          torch.randn(96, device=DEV))
    g96, b96 = torch.ones(96, device=DEV), torch.zeros(96, device=DEV)
    report("evo_se_affine 96ch", time_it(lambda: ops.evo_se_affine(st96, g96, b96, ab96[0], ab96[1], 64 ** 3,
                                                                     chan_sum=cs, se=se)), 1)
    # blend of one window batch
    logits = torch.randn((n, 3, 128, 128, 128), device=DEV)
    acc = torch.zeros((3, 240, 240, 160), device=DEV)
    prof = [torch.ones(128, device=DEV)] * 3
    org = [(0, 0, 0), (0, 0, 32), (0, 96, 0), (0, 96, 32)]
    report("blend_accumulate 4 windows", time_it(lambda: ops.blend_accumulate(logits, acc, prof, org)),
           logits.numel() * 4 * 3)


if __name__ == "__main__":
    main()
