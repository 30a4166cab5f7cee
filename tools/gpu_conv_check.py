"""GPU bring-up check for the conv kernels: compares against torch F.conv3d (fp32, TF32 off) on identical
bf16-rounded inputs, and times the L1/L2 shapes that hold most of the FLOPs.  Run under gpurun:
    python tools/gpu_conv_check.py [--perf]
Writes a JSON summary to gpurun_out/conv_check.json.
"""
import json
import os
import sys
import time
import traceback

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from brats21_b200 import ops  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def ref_conv(x_cl, w, b, dil):
    x = x_cl.float().permute(0, 4, 1, 2, 3).contiguous()
    k = w.shape[2]
    wq = w.to(torch.bfloat16).float()
    y = F.conv3d(x, wq, b, padding=dil if k == 3 else 0, dilation=dil)
    return y.permute(0, 2, 3, 4, 1).contiguous()


def one_case(name, n, d, h, w, cin, cout, k, dil, bias=True, stats=True, ldx=None, ldy=None, conv=ops.conv3d):
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(hash(name) % (2 ** 31))
    cin_p = (cin + 7) // 8 * 8
    xb = torch.zeros((n, d, h, w, ldx or cin_p), dtype=torch.bfloat16, device=dev)
    xb[..., :cin] = torch.randn((n, d, h, w, cin), device=dev, generator=g).to(torch.bfloat16)
    x = xb[..., :cin_p]
    wt = torch.randn((cout, cin, k, k, k), device=dev, generator=g) / (cin * k ** 3) ** 0.5
    b = torch.randn((cout,), device=dev, generator=g) if bias else None
    pw = ops.PackedConv(wt, b, cin_padded=cin_p)
    yb = torch.full((n, d, h, w, ldy or cout), 7.0, dtype=torch.bfloat16, device=dev)
    y = yb[..., :cout]
    st = ops.new_stats(n, dev) if stats else None
    conv(x, pw, out=y, stats=st, dil=dil)
    torch.cuda.synchronize()
    ref = ref_conv(x[..., :cin], wt, b, dil)
    err = (y.float() - ref).abs().max().item()
    scale = ref.abs().max().item()
    res = {"case": name, "max_abs_err": err, "ref_max": scale, "ok": bool(err <= 2e-2 * max(scale, 1.0))}
    if ldy and ldy > cout:
        res["pad_untouched"] = bool((yb[..., cout:] == 7.0).all().item())
        res["ok"] = res["ok"] and res["pad_untouched"]
    if stats:
        s = st.sum(0)  # [n, 8, 2]
        gs = cout // 8
        r = ref.reshape(n, -1, 8, gs).double()
        rs = r.sum(dim=(1, 3))
        rq = (r * r).sum(dim=(1, 3))
        e1 = ((s[..., 0] - rs).abs() / (rs.abs() + 1.0)).max().item()
        e2 = ((s[..., 1] - rq).abs() / (rq.abs() + 1.0)).max().item()
        res["stats_rel_err"] = max(e1, e2)
        res["ok"] = res["ok"] and res["stats_rel_err"] < 1e-3
    return res


def perf_case(name, n, s, cin, cout, k, dil, iters=10, conv=ops.conv3d):
    dev = "cuda"
    x = torch.randn((n, s, s, s, cin), device=dev).to(torch.bfloat16)
    wt = torch.randn((cout, cin, k, k, k), device=dev) / (cin * k ** 3) ** 0.5
    pw = ops.PackedConv(wt, torch.zeros(cout, device=dev))
    y = torch.empty((n, s, s, s, cout), dtype=torch.bfloat16, device=dev)
    st = ops.new_stats(n, dev)
    for _ in range(3):
        conv(x, pw, out=y, stats=st, dil=dil)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        conv(x, pw, out=y, stats=st, dil=dil)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    flops = 2.0 * n * s ** 3 * cin * cout * k ** 3
    # cuDNN bf16 channels-last for comparison
    xc = x.permute(0, 4, 1, 2, 3)
    wc = wt.to(torch.bfloat16).to(memory_format=torch.channels_last_3d)
    for _ in range(3):
        F.conv3d(xc, wc, None, padding=dil if k == 3 else 0, dilation=dil)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        F.conv3d(xc, wc, None, padding=dil if k == 3 else 0, dilation=dil)
    e1.record()
    torch.cuda.synchronize()
    ms_cudnn = e0.elapsed_time(e1) / iters
    return {"case": name, "ms": ms, "tflops": flops / ms / 1e9, "cudnn_ms": ms_cudnn,
            "cudnn_tflops": flops / ms_cudnn / 1e9}


def main():
    out = {"correctness": [], "perf": []}
    cases = [
        # name, n, d, h, w, cin, cout, k, dil
        ("c48_48", 1, 16, 16, 16, 48, 48, 3, 1),
        ("c4_48_first", 1, 16, 16, 16, 4, 48, 3, 1),
        ("c96_96", 2, 8, 16, 16, 96, 96, 3, 1),
        ("c192_192", 1, 8, 8, 8, 192, 192, 3, 1),
        ("c384_384", 1, 8, 8, 8, 384, 384, 3, 1),
        ("c768_192", 1, 8, 8, 8, 768, 192, 3, 1),
        ("c384_96_d2", 1, 16, 16, 16, 384, 96, 3, 2),
        ("c384_96_d4", 1, 16, 16, 16, 384, 96, 3, 4),
        ("c384_96_d6", 1, 16, 16, 16, 384, 96, 3, 6),
        ("k1_48_24", 1, 16, 16, 16, 48, 24, 1, 1),
        ("k1_384_384", 1, 8, 8, 8, 384, 384, 1, 1),
        ("odd_shape", 1, 10, 12, 20, 48, 48, 3, 1),
        ("odd_shape2", 2, 5, 7, 9, 16, 16, 3, 1),
        ("tiny4", 1, 4, 4, 4, 128, 128, 3, 1),
        ("tiny2", 1, 2, 2, 2, 64, 64, 3, 2),
        ("w16", 1, 32, 32, 32, 16, 16, 3, 1),
        # plane-marching kernel: several d segments / work items per CTA, ragged tiles, ring wrap-around
        ("m48_long", 2, 40, 24, 24, 48, 48, 3, 1),
        ("m48_ragged", 1, 21, 19, 13, 48, 48, 3, 1),
        ("m32_32", 1, 20, 16, 32, 32, 32, 3, 1),
        ("m16_64", 1, 9, 32, 16, 16, 64, 3, 1),
        ("m8_48_d1", 3, 1, 16, 8, 8, 48, 3, 1),
        ("m48_many", 1, 24, 160, 160, 48, 48, 3, 1),
    ]
    if "--march-only" in sys.argv:
        cases = [c for c in cases if c[7] == 3 and c[8] == 1 and c[6] in (16, 32, 48, 64) and c[5] <= 48]
    for c in cases:
        try:
            r = one_case(*c)
            r["kernel"] = "march" if (ops.use_march and c[7] == 3 and c[8] == 1 and c[6] in (16, 32, 48, 64)
                                      and c[5] <= 48 and c[3] >= 8 and c[4] >= 8) else "tap"
        except Exception as e:  # noqa: BLE001
            r = {"case": c[0], "ok": False, "error": repr(e)[:160]}
        print(r, flush=True)
        out["correctness"].append(r)
        if "error" in r and "CUDA error" in r["error"]:
            break  # the context is gone; later cases would only repeat the error
    try:
        r = one_case("slices", 1, 8, 8, 8, 48, 24, 1, 1, ldx=96, ldy=48)
    except Exception as e:  # noqa: BLE001
        r = {"case": "slices", "ok": False, "error": repr(e)[:160]}
    print(r, flush=True)
    out["correctness"].append(r)

    if "--perf1" in sys.argv:
        print(perf_case("L1_48_48_b4", 4, 128, 48, 48, 3, 1), flush=True)
        return
    if "--perf" in sys.argv:
        perf = [
            ("L1_48_48_b4", 4, 128, 48, 48, 3, 1),
            ("L1_8_48_b4", 4, 128, 8, 48, 3, 1),
            ("L1_48_48_b1", 1, 128, 48, 48, 3, 1),
            ("L1_96_48_b4", 4, 128, 96, 48, 3, 1),
            ("L2_96_96_b4", 4, 64, 96, 96, 3, 1),
            ("L3_192_192_b4", 4, 32, 192, 192, 3, 1),
            ("L4_384_384_b4", 4, 16, 384, 384, 3, 1),
            ("L1_k1_48_24_b4", 4, 128, 48, 24, 1, 1),
        ]
        for c in perf:
            try:
                r = perf_case(*c)
                if c[5] == 3 and c[4] in (16, 32, 48, 64) and c[3] <= 48:  # also time the tap kernel on these
                    ops.use_march = False
                    r["tap_ms"] = perf_case(*c)["ms"]
                    ops.use_march = True
            except Exception as e:  # noqa: BLE001
                r = {"case": c[0], "error": repr(e)[:160]}
            print(r, flush=True)
            out["perf"].append(r)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/conv_check.json", "w") as f:
        json.dump(out, f, indent=1)
    print("ALL_OK" if all(r.get("ok") for r in out["correctness"]) else "SOME_FAILED")


if __name__ == "__main__":
    if "--no-march" in sys.argv:
        ops.use_march = False
    t0 = time.time()
    main()
    print("elapsed", time.time() - t0)
