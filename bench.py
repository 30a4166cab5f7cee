#!/usr/bin/env python
"""bench.py — headline benchmark of the BraTS21 hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload NAME]

One *step* = one full pass of the hot path over one synthetic 4x240x240x155 volume:
  workload "v2_tta8" (default; BASELINE.json configs[2], the configuration the volumes/s metric is quoted on):
      EquiUNet-ASPP-Evo (width 48), 8 axis-flip TTA variants, 128^3 sliding window (overlap 0.25 -> 18 windows per
      variant, 144 per volume, batches of 4), gaussian blending, sigmoid/mean/threshold, BraTS label map.
  workload "v1_sw"  (configs[1]): EquiUNet V1, no TTA, same window grid, batches of 4.
  workload "v2_ens3_tta8" (configs[4]): three EquiUNet-ASPP-Evo models x 8 flips per volume (432 windows), cohort
      volumes sharded across ranks.   workload "v2_train" (configs[3]): one training step, batch 1 per GPU.
`value` is whole-job volumes/s with the volume resident in HBM; `e2e` is the same through the public API
(brats21_b200.engine.predict_volume) from a pinned HOST volume to HOST uint8 labels, copies inside the timed region.
Multi-GPU (torchrun): volumes are sharded across ranks, no data-path collective ("weak" scaling).
`--impl reference` times the reference's own CPU implementation of the path (its torch-fp32 restatement in oracle/,
the reference being pure Python that cannot travel to the GPU box) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

# stdout carries exactly ONE JSON line: NCCL writes its "NCCL version ..." banner (any NCCL_DEBUG level >= VERSION) and
# its debug log to stdout unless told otherwise, so send them to stderr
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":  # that level exists only to print the banner
    os.environ.pop("NCCL_DEBUG")

import torch  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

VOL_SHAPE = (240, 240, 155)
ROI = (128, 128, 128)
WIDTH = 48
WORKLOADS = {
    "v2_tta8": dict(version=2, tta="flip8", mode="gaussian", sw_batch=4, seed=93,
                    desc="EquiUNet-ASPP-Evo w48, 8-flip TTA, 128^3 sliding window (144 windows), gaussian blend, "
                         "labels; one synthetic 4x240x240x155 volume per step"),
    "v2_ens3_tta8": dict(version=2, tta="flip8", mode="gaussian", sw_batch=4, seed=93, ensemble=(93, 123, 7),
                         desc="Model-6-style ensemble of 3 EquiUNet-ASPP-Evo w48 x 8-flip TTA (432 windows per volume), "
                              "gaussian blend, mean over 24 probability maps, labels; cohort volumes sharded across ranks"),
    "v1_sw": dict(version=1, tta=None, mode="constant", sw_batch=4, seed=123,
                  desc="EquiUNet V1 w48, no TTA, 128^3 sliding window (18 windows, batches of 4), labels"),
    "v2_train": dict(version=2, train=True, seed=93,
                     desc="EquiUNet-ASPP-Evo w48 training step: forward, Dice over 3 heads, backward, fused Ranger; "
                          "one synthetic 4x128^3 crop per GPU per step (batch 1/GPU), data-parallel over ranks"),
}


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7),
                              ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower() == "active":
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1590.0)), d.get("hbm_gbs", 6650.0), "measured"
    return 1590.0, 6650.0, "fallback"


def pad_to_8(x):
    """shape_to_divisible(k=8) (utils/transforms.py:483-512) on the last three dims."""
    import torch.nn.functional as F
    pads, meta = [], []
    for s in reversed(x.shape[-3:]):
        p = (-s) % 8
        pb = (p + 1) // 2
        pads += [pb, p - pb]
        meta.append((pb, s))
    return (F.pad(x, pads) if any(pads) else x), list(reversed(meta))


# ------------------------------------------------------------------------------------------------ CPU baseline
def cpu_window_forward_seconds(version: int, repeats: int = 1, shape=ROI):
    """The reference's CPU path for one window: torch fp32 on all host cores (oracle port of the networks)."""
    from oracle import nets, synth  # CHECKER/BASELINE use of oracle/ (allowed for the cpu_baseline leg only)
    torch.set_num_threads(os.cpu_count() or 1)
    params = synth.make_params(version, WIDTH, 93 if version == 2 else 123)
    fwd = nets.equiunet_v2_forward if version == 2 else nets.equiunet_v1_forward
    x = synth.volume(seed=0, shape=shape)
    times = []
    with torch.no_grad():
        for _ in range(repeats):
            t0 = time.perf_counter()
            fwd(params, x, deep_supervision=False)
            times.append(time.perf_counter() - t0)
    return times


def torch_gpu_window_seconds(version: int, dev, batch: int = 4):
    """Context for the headline: the reference's network code (oracle port, torch/cuDNN) on the SAME GPU
    under torch.autocast(bf16) — its default mixed-precision mode — for one batch of 128^3 windows.  Baseline only."""
    from oracle import nets, synth  # CHECKER/BASELINE use of oracle/ (part of the cpu_baseline leg)
    params = {k: v.to(dev) for k, v in synth.make_params(version, WIDTH, 93 if version == 2 else 123).items()}
    fwd = nets.equiunet_v2_forward if version == 2 else nets.equiunet_v1_forward
    x = torch.cat([synth.volume(seed=s, shape=ROI) for s in range(batch)]).to(dev)
    best = None
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        for _ in range(3):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fwd(params, x, deep_supervision=False)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            best = ms if best is None else min(best, ms)
    del params, x
    torch.cuda.empty_cache()
    return best * 1e-3 / batch


def windows_per_volume(wl):
    return 18 * (8 if wl["tta"] == "flip8" else 1) * len(wl.get("ensemble", (0,)))


def run_reference(args, wl, rank, world):
    """`--impl reference`: rank 0 only; each step = one 128^3 window through the CPU path; volumes/s extrapolated
    over the 144 (or 18) windows of the workload (blend/TTA arithmetic is <1% of the CPU time)."""
    if rank != 0:
        return
    nwin = windows_per_volume(wl)
    shape = ROI
    t_probe = cpu_window_forward_seconds(wl["version"], 1, shape)[0]  # first (untimed) warm-up step
    scale = 1.0
    if t_probe * (args.steps + args.warmup) > 240.0:  # keep the whole run within a few minutes
        shape, scale = (64, 64, 64), 8.0
    for _ in range(max(args.warmup - 1, 0)):
        cpu_window_forward_seconds(wl["version"], 1, shape)
    times = cpu_window_forward_seconds(wl["version"], args.steps, shape)
    t_win = statistics.mean(times) * scale
    value = 1.0 / (nwin * t_win)
    sample = f"1 of {nwin} windows per step (V{wl['version']} forward, {shape[0]}^3 fp32" + \
        (", x8 voxel scaling to 128^3" if scale != 1.0 else "") + "), extrapolated to the volume"
    line = {"impl": "reference", "metric": "volumes/s (sliding-window + 8xTTA)", "value": value, "unit": "volumes/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / value,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "description": wl["desc"]},
            "cpu_baseline": {"value": value, "unit": "volumes/s", "cores": os.cpu_count(), "kind": "port",
                             "sample": sample},
            "e2e": {"value": value, "unit": "volumes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ B200 arm
def run_b200_train(args, wl, rank, local_rank, world):
    """BASELINE configs[3]: V2 training step, batch 1 per GPU, data-parallel (NCCL all-reduce overlapped with backward)."""
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import warnings
    from brats21_b200 import _lib, engine, networks, ops, parallel, synth
    from brats21_b200.losses import DiceLoss
    from brats21_b200.optimizer import Ranger2020
    torch.manual_seed(wl["seed"])
    feats = [WIDTH * 2 ** i for i in range(4)]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        net = networks.EquiUnetASSPEvo(4, 3, feats, norm_layer="group", act="relu", deep_supervision=True).to(dev).train()
    model = parallel.DistributedDataParallel(net) if world > 1 else net
    opt = Ranger2020([p for n, p in net.named_parameters() if not n.endswith(".v")], lr=3e-4, weight_decay=1e-5)
    if world > 1:
        model.attach_optimizer(opt)
    crit = DiceLoss()
    host_img = synth.volume(seed=2000 + rank, shape=ROI).pin_memory()
    host_tgt = synth.target(shape=ROI).pin_memory()
    host_loss = torch.empty((1,), dtype=torch.float32).pin_memory()
    img_dev, tgt_dev = host_img.to(dev), host_tgt.to(dev)

    def step_device():
        return engine.train_step(None, model, crit, opt, img_dev, tgt_dev)

    def step_e2e():
        img = host_img.to(dev, non_blocking=True)
        tgt = host_tgt.to(dev, non_blocking=True)
        loss = engine.train_step(None, model, crit, opt, img, tgt)
        host_loss.copy_(loss.reshape(1), non_blocking=True)
        torch.cuda.current_stream().synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms

    for _ in range(args.warmup):
        step_device()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = _lib.launch_count
    ms = timed(step_device, args.steps)
    launches = _lib.launch_count - l0
    clocks = sampler.stop()
    step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    ops.conv_profile = []
    step_device()
    torch.cuda.synchronize()
    prof, ops.conv_profile = ops.conv_profile, None
    conv_ms = sum(a.elapsed_time(b) for a, b, _, _ in prof)
    conv_flops = sum(f for _, _, f, _ in prof)
    peak_tf, _, peak_kind = measured_peaks()
    achieved = conv_flops / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
    value = world * args.steps / (ms * 1e-3)
    line = {"metric": "train patches/s (EquiUNet-ASPP-Evo, 128^3, batch 1/GPU)", "value": value, "unit": "patches/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": args.workload, "description": wl["desc"],
                       "l2": "activations of one step (>3 GB) far exceed the 126 MB L2",
                       "parallelism": f"dp{world}: flat-buffer bucketed NCCL all-reduce overlapped with backward"},
            "clocks": clocks,
            "e2e": {"value": world * args.steps / (ms_e2e * 1e-3), "unit": "patches/s", "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": (host_img.numel() + host_tgt.numel()) * 4, "d2h_bytes_per_step": 4},
            "gpu_launches": launches,
            "roofline": {"bound": "tensor", "kernel": "conv3d fwd + dgrad + wgrad (tcgen05), all launches of one step",
                         "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
                         "peak_kind": f"{peak_kind} bf16 (sustained)", "traffic": None, "launches": len(prof),
                         "flops_per_launch": conv_flops / max(len(prof), 1), "avg_launch_ms": conv_ms / max(len(prof), 1),
                         "share_of_step": conv_ms / (ms / args.steps) if ms > 0 else None}}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_b200(args, wl, rank, local_rank, world):
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (no CPU fallback for the product path)")
    if wl.get("train"):
        return run_b200_train(args, wl, rank, local_rank, world)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    from brats21_b200 import _lib, engine, networks, ops, synth, tta

    feats = [WIDTH * 2 ** i for i in range(4)]
    cls = networks.EquiUnetASSPEvo if wl["version"] == 2 else networks.EquiUnet
    import warnings
    nets = []
    for seed in wl.get("ensemble", (wl["seed"],)):  # the reference's seeds (arguments_train.py:103) + 7
        torch.manual_seed(seed)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            nets.append(cls(4, 3, feats, norm_layer="group", act="relu", deep_supervision=True).to(dev).eval())
    comp = tta.get_flip8_transforms() if wl["tta"] == "flip8" else None

    host_vol = synth.volume(seed=1000 + rank, shape=VOL_SHAPE).pin_memory()
    host_lab = torch.empty((1, 1) + VOL_SHAPE, dtype=torch.uint8).pin_memory()
    vol_dev, _ = pad_to_8(host_vol.to(dev))

    def step_device():
        return engine.predict_volume(nets, vol_dev, comp, True, ROI, wl["sw_batch"], 0.25, wl["mode"])

    def step_e2e():
        v, meta = pad_to_8(host_vol.to(dev, non_blocking=True))
        _, label = engine.predict_volume(nets, v, comp, True, ROI, wl["sw_batch"], 0.25, wl["mode"])
        crop = label[..., meta[0][0]:meta[0][0] + meta[0][1], meta[1][0]:meta[1][0] + meta[1][1],
                     meta[2][0]:meta[2][0] + meta[2][1]]
        host_lab.copy_(crop, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            import torch.distributed as dist
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms

    for _ in range(args.warmup):
        step_device()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = _lib.launch_count
    ms = timed(step_device, args.steps)
    launches = _lib.launch_count - l0
    clocks = sampler.stop()
    step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    # one more step, launched eagerly with a CUDA-event pair around every conv launch (the timed steps above replay
    # the per-batch forward from a CUDA graph): live per-launch durations of the dominant kernel for the roofline
    ops.conv_profile = []
    step_device()
    torch.cuda.synchronize()
    prof, ops.conv_profile = ops.conv_profile, None

    # dominant kernel (conv implicit GEMM): algorithmic FLOPs / event-timed launch durations
    conv_ms = sum(a.elapsed_time(b) for a, b, _, _ in prof)
    conv_flops = sum(f for _, _, f, _ in prof)
    peak_tf, _, peak_kind = measured_peaks()
    achieved = conv_flops / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
    by_kind = {}
    for a, b, fl, key in prof:
        r = by_kind.setdefault(str(key[0]), [0, 0.0, 0.0])
        r[0] += 1
        r[1] += a.elapsed_time(b)
        r[2] += fl
    by_kind = {k: {"launches": v[0], "ms": v[1], "tflops": v[2] / (v[1] * 1e-3) / 1e12 if v[1] > 0 else 0.0}
               for k, v in by_kind.items()}
    roofline = {"bound": "tensor", "kernel": "conv3d implicit GEMM (tcgen05): all conv launches of one step "
                "(march = plane-marching 48-ch layers, slide = sliding-window 96-ch layers, tap = generic, point = 1x1)",
                "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
                "peak_kind": f"{peak_kind} bf16 (sustained)", "by_kernel": by_kind,
                # DRAM bytes per launch of the dominant kernel (conv_march_kernel<48>, 48->48 at 4x128^3) from the
                # ncu --set full capture profiles/r01c_conv_full.md; algorithmic bytes = 2 * 4*128^3 * 48 * 2 B = 1.61e9
                "traffic": 1.577e9, "traffic_kernel": "conv_march_kernel<48> 48->48 @4x128^3",
                "traffic_algorithmic": 1.611e9,
                "launches": len(prof), "flops_per_launch": conv_flops / max(len(prof), 1),
                "avg_launch_ms": conv_ms / max(len(prof), 1),
                "share_of_step": conv_ms / (ms / args.steps) if ms > 0 else None}

    value = world * args.steps / (ms * 1e-3)
    e2e_value = world * args.steps / (ms_e2e * 1e-3)
    line = {"metric": "volumes/s (sliding-window + 8xTTA)", "value": value, "unit": "volumes/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": args.workload, "description": wl["desc"], "roi": list(ROI), "overlap": 0.25,
                       "sw_batch_size": wl["sw_batch"], "windows_per_volume": windows_per_volume(wl),
                       "l2": "per-step working set (>4 GB of activations per window batch) far exceeds the 126 MB L2",
                       "sharding": "volumes across ranks, no collective"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "volumes/s", "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": host_vol.numel() * 4, "d2h_bytes_per_step": host_lab.numel()},
            "gpu_launches": launches, "roofline": roofline}
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            nwin = windows_per_volume(wl)
            t = cpu_window_forward_seconds(wl["version"], 1)[0]
            line["cpu_baseline"] = {"value": 1.0 / (nwin * t), "unit": "volumes/s", "cores": os.cpu_count(),
                                    "kind": "port", "sample": f"1 of {nwin} windows (V{wl['version']} forward 128^3 "
                                    f"fp32, {t:.1f} s), extrapolated to the volume"}
            try:
                tw = torch_gpu_window_seconds(wl["version"], dev)
                line["torch_gpu_baseline"] = {
                    "value": 1.0 / (nwin * tw), "unit": "volumes/s", "kind": "port",
                    "sample": f"reference network code (torch/cuDNN, autocast bf16) on this GPU: best of 3 batches of 4 "
                              f"windows, {tw * 1e3:.1f} ms per window, network forward only, extrapolated to {nwin} windows"}
            except Exception as exc:  # noqa: BLE001  (a baseline must never break the bench line)
                line["torch_gpu_baseline"] = {"unavailable": repr(exc)[:200]}
        print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="v2_tta8", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank, local_rank, world = dist_env()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl, rank, world)
    else:
        run_b200(args, wl, rank, local_rank, world)


if __name__ == "__main__":
    main()
