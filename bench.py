#!/usr/bin/env python
"""bench.py — headline benchmark of the BraTS21 hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload NAME]

One *step* = one full pass of the hot path over one synthetic 4x240x240x155 volume:
  workload "v2_tta8" (default; BASELINE.json configs[2], the configuration the volumes/s metric is quoted on):
      EquiUNet-ASPP-Evo (width 48), 8 axis-flip TTA variants, 128^3 sliding window (overlap 0.25 -> 18 windows per
      variant, 144 per volume, batches of 9), gaussian blending, sigmoid/mean/threshold, BraTS label map.
  workload "v1_sw"  (configs[1]): EquiUNet V1, no TTA, same window grid, batches of 4.
  workload "v2_ens3_tta8" (configs[4]): three EquiUNet-ASPP-Evo models x 8 flips per volume (432 windows), cohort
      volumes sharded across ranks.   workload "v2_train" (configs[3]): one training step, batch 1 per GPU.
`value` is whole-job volumes/s with the volume resident in HBM; `e2e` is the same through the public API
(brats21_b200.engine.predict_volume) from a pinned HOST volume to HOST uint8 labels, copies inside the timed region.
Multi-GPU (torchrun): volumes are sharded across ranks, no data-path collective ("weak" scaling).
The default line also carries a `train` block — K' steps of workload "v2_train" at the same N ranks (data-parallel over
NCCL: patches/s, ms/step, e2e, conv roofline, all-reduce bytes, exposed communication time) — so that the driver's
BENCH / SCALE records hold both halves of BASELINE.json's metric; an `hbm` block (event-timed GB/s of the
bandwidth-bound kernels against the measured HBM peak); and a `parity` block quoting the full-size parity record
(profiles/r02_parity_full_size.json, written by tests/test_gpu_fullsize.py on a B200).
Workload "v2_ens3_tta8_cohort" (configs[4]) runs --cohort synthetic volumes (default 219) sharded over the ranks.
`--impl reference` times the reference's own CPU implementation of the path — the UNMODIFIED reference network
(oracle/_ref, built by oracle/make_ref.py; the torch-fp32 restatement in oracle/ when that archive is absent) — on the
box's host cores, on a bounded sample of the same workload.
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
    # sw_batch_size is a free knob of sliding_window_inference (results do not depend on it: per-sample statistics,
    # window-ordered blending); BASELINE configs[2] does not fix it: 18 windows per variant = 2 batches of 9
    "v2_tta8": dict(version=2, tta="flip8", mode="gaussian", sw_batch=9, seed=93,
                    desc="EquiUNet-ASPP-Evo w48, 8-flip TTA, 128^3 sliding window (144 windows, batches of 9), gaussian "
                         "blend, labels; one synthetic 4x240x240x155 volume per step"),
    "v2_ens3_tta8": dict(version=2, tta="flip8", mode="gaussian", sw_batch=9, seed=93, ensemble=(93, 123, 7),
                         desc="Model-6-style ensemble of 3 EquiUNet-ASPP-Evo w48 x 8-flip TTA (432 windows per volume), "
                              "gaussian blend, mean over 24 probability maps, labels; cohort volumes sharded across ranks"),
    "v2_ens3_tta8_cohort": dict(version=2, tta="flip8", mode="gaussian", sw_batch=9, seed=93, ensemble=(93, 123, 7),
                                cohort=True,
                                desc="BASELINE configs[4]: 3-model EquiUNet-ASPP-Evo w48 ensemble x 8-flip TTA over a cohort of "
                                     "synthetic 4x240x240x155 volumes (seeds 0..n-1) sharded round-robin across ranks; host "
                                     "volume in, host labels out, prefetch thread generating the next volume"),
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
def _cpu_forward_fn(version: int):
    """(callable x -> logits, kind): the UNMODIFIED reference network module on the host cores when oracle/_ref (or
    /root/reference) is present — kind "reference" — else the oracle's torch-fp32 restatement — kind "port"."""
    from oracle import nets, ref_loader, synth  # CHECKER/BASELINE use of oracle/ (cpu_baseline / --impl reference legs)
    params = synth.make_params(version, WIDTH, 93 if version == 2 else 123)
    if ref_loader.available():
        try:
            ref = ref_loader.load()
            feats = [WIDTH * 2 ** i for i in range(4)]
            with ref_loader.quiet():
                cls = ref.equiunet2021.EquiUnetASSPEvo if version == 2 else ref.equiunet2020.EquiUnet
                net = cls(4, 3, feats, norm_layer="group", act="relu", deep_supervision=False)
            net.load_state_dict({k: v for k, v in params.items() if not k.startswith("deep")}, strict=True)
            net.eval()
            return (lambda x: net(x)), "reference"
        except Exception as exc:  # noqa: BLE001  (fall back to the port, say so)
            print(f"bench.py: unmodified reference unavailable ({exc!r}); timing the oracle port", file=sys.stderr)
    fwd = nets.equiunet_v2_forward if version == 2 else nets.equiunet_v1_forward
    return (lambda x: fwd(params, x, deep_supervision=False)), "port"


def cpu_window_forward_seconds(version: int, repeats: int = 1, shape=ROI, fn=None):
    """The reference's CPU path for one window: torch fp32 on all host cores."""
    from oracle import synth
    torch.set_num_threads(os.cpu_count() or 1)
    kind = None
    if fn is None:
        fn, kind = _cpu_forward_fn(version)
    x = synth.volume(seed=0, shape=shape)
    times = []
    with torch.no_grad():
        for _ in range(repeats):
            t0 = time.perf_counter()
            fn(x)
            times.append(time.perf_counter() - t0)
    return times, kind


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
    """`--impl reference`: rank 0 only; each step = one 128^3 window through the reference's CPU path (the unmodified
    reference network when oracle/_ref is present); volumes/s extrapolated over the 144 (or 18) windows of the
    workload (blend/TTA arithmetic is <1% of the CPU time).  Training workload: patches/s from forward time x 3."""
    if rank != 0:
        return
    train = bool(wl.get("train"))
    nwin = 1 if train else windows_per_volume(wl)
    shape = ROI
    fn, kind = _cpu_forward_fn(wl["version"])
    t_probe = cpu_window_forward_seconds(wl["version"], 1, shape, fn)[0][0]  # first (untimed) warm-up step
    scale = 1.0
    if t_probe * (args.steps + args.warmup) > 240.0:  # keep the whole run within a few minutes
        shape, scale = (64, 64, 64), 8.0
    for _ in range(max(args.warmup - 1, 0)):
        cpu_window_forward_seconds(wl["version"], 1, shape, fn)
    times, _ = cpu_window_forward_seconds(wl["version"], args.steps, shape, fn)
    t_win = statistics.mean(times) * scale
    if train:
        value, unit, metric = 1.0 / (3.0 * t_win), "patches/s", "train patches/s (EquiUNet-ASPP-Evo, 128^3, batch 1/GPU)"
        sample = f"forward of one {shape[0]}^3 crop per step (fp32, all host cores); step = 3 x forward (fwd + dgrad + wgrad)"
    else:
        value, unit, metric = 1.0 / (nwin * t_win), "volumes/s", "volumes/s (sliding-window + 8xTTA)"
        sample = f"1 of {nwin} windows per step (V{wl['version']} forward, {shape[0]}^3 fp32" + \
            (", x8 voxel scaling to 128^3" if scale != 1.0 else "") + "), extrapolated to the volume"
    line = {"impl": "reference", "metric": metric, "value": value, "unit": unit,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / value,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, wl),
            "cpu_baseline": {"value": value, "unit": unit, "cores": os.cpu_count(), "kind": kind,
                             "sample": sample},
            "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(args, wl):
    """The `config` object: identical for the b200 and the reference arm of one workload."""
    cfg = {"workload": args.workload, "description": wl["desc"]}
    if not wl.get("train"):
        cfg.update({"roi": list(ROI), "overlap": 0.25, "sw_batch_size": wl["sw_batch"],
                    "windows_per_volume": windows_per_volume(wl),
                    "l2": "per-step working set (>4 GB of activations per window batch) far exceeds the 126 MB L2",
                    "sharding": "volumes across ranks, no collective"})
        if wl.get("cohort"):
            cfg["cohort_volumes"] = args.cohort
    else:
        cfg.update({"l2": "activations of one step (>3 GB) far exceed the 126 MB L2",
                    "parallelism": "data-parallel: flat-buffer bucketed NCCL all-reduce overlapped with backward"})
    return cfg


# ------------------------------------------------------------------------------------------------ B200 arm
class Dist:
    """One process per GPU; NCCL process group when launched under torchrun."""

    def __init__(self):
        self.rank, self.local_rank, self.world = dist_env()
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def timed(self, fn, steps):
        """CUDA-event time of `steps` calls, bracketed by barrier + synchronize, MAX over ranks (ms)."""
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        self.barrier()
        return self.max_over_ranks(e0.elapsed_time(e1))

    def max_over_ranks(self, v: float) -> float:
        if self.world > 1:
            import torch.distributed as dist
            t = torch.tensor([v], device=self.dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return t.item()
        return v

    def close(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


def summarize_profile(prof, per_key=lambda key: str(key)):
    by = {}
    for a, b, amount, key in prof:
        r = by.setdefault(per_key(key), [0, 0.0, 0.0])
        r[0] += 1
        r[1] += a.elapsed_time(b)
        r[2] += amount
    return by


def measure_train(args, dx: Dist, steps: int, warmup: int, seed: int = 93):
    """BASELINE configs[3]: V2 training step (forward, Dice over 3 heads, hand-scheduled backward, fused Ranger incl. the
    bf16 weight re-pack), batch 1 per GPU, data-parallel (NCCL all-reduce of the flat fp32 gradient buffer in buckets on
    a side stream, overlapped with the backward).  Returns the `train` block."""
    import warnings
    from brats21_b200 import _lib, engine, networks, ops, parallel, synth
    from brats21_b200.losses import DiceLoss
    from brats21_b200.optimizer import Ranger2020
    dev, world, rank = dx.dev, dx.world, dx.rank
    torch.manual_seed(seed)
    feats = [WIDTH * 2 ** i for i in range(4)]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        net = networks.EquiUnetASSPEvo(4, 3, feats, norm_layer="group", act="relu", deep_supervision=True).to(dev).train()
    model = parallel.DistributedDataParallel(net) if world > 1 else net
    opt = Ranger2020([p for n, p in net.named_parameters() if not n.endswith(".v")], lr=3e-4, weight_decay=1e-5,
                     use_gc=False)
    if world > 1:
        model.attach_optimizer(opt)
    crit = DiceLoss()
    host_img = synth.volume(seed=2000 + rank, shape=ROI).pin_memory()
    host_tgt = synth.target(shape=ROI).pin_memory()
    host_loss = torch.empty((1,), dtype=torch.float32).pin_memory()
    img_dev, tgt_dev = host_img.to(dev), host_tgt.to(dev)

    # the public API of the step: engine.TrainStep = engine.train_step replayed from one CUDA graph per input shape
    graphed = engine.TrainStep(model, crit, opt)

    def step_eager():
        opt.disable_graph_mode()
        return engine.train_step(None, model, crit, opt, img_dev, tgt_dev)

    def step_device():
        return graphed(img_dev, tgt_dev)

    # e2e: every step copies ITS batch from pinned host memory and reads its loss back; as a DataLoader with
    # pin_memory + non_blocking does, the copy of the next batch runs on a copy stream under the current step
    copy_stream = torch.cuda.Stream(device=dev)
    staged = [[torch.empty_like(img_dev), torch.empty_like(tgt_dev), None] for _ in range(2)]
    state = {"i": 0}

    def stage(slot):
        with torch.cuda.stream(copy_stream):
            if slot[2] is not None:
                copy_stream.wait_event(slot[2])  # the step that last read this slot has finished with it
            slot[0].copy_(host_img, non_blocking=True)
            slot[1].copy_(host_tgt, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return ev

    ready = [stage(staged[0]), None]

    def step_e2e():
        i = state["i"]
        cur, nxt = staged[i % 2], staged[(i + 1) % 2]
        ready[(i + 1) % 2] = stage(nxt)                      # prefetch the next batch
        torch.cuda.current_stream().wait_event(ready[i % 2])
        loss = graphed(cur[0], cur[1])
        done = torch.cuda.Event()
        done.record()
        cur[2] = done
        host_loss.copy_(loss.reshape(1), non_blocking=True)
        torch.cuda.current_stream().synchronize()
        state["i"] = i + 1

    for _ in range(max(warmup, graphed.eager_warmup + 1)):
        step_device()
    l0 = _lib.launch_count
    ms = dx.timed(step_device, steps)
    launches = _lib.launch_count - l0
    step_e2e()
    ms_e2e = dx.timed(step_e2e, steps)
    loss_value = float(host_loss.item())
    comm = None
    if world > 1:
        # exposed (non-overlapped) communication: the same steps with the bucket all-reduces switched off
        gs = net.grad_store()
        hooks = (gs.on_begin, gs.on_ready, gs.on_finish)
        nbuckets, grad_bytes = len(model._reducer.bounds), gs.flat.numel() * 4
        gs.on_begin = gs.on_ready = gs.on_finish = None
        nocomm = engine.TrainStep(model, crit, opt)
        for _ in range(nocomm.eager_warmup + 1):
            nocomm(img_dev, tgt_dev)
        ms_nocomm = dx.timed(lambda: nocomm(img_dev, tgt_dev), steps)
        gs.on_begin, gs.on_ready, gs.on_finish = hooks
        comm = {"allreduce_bytes_per_step": grad_bytes, "buckets": nbuckets, "backend": "nccl",
                "ms_per_step_without_allreduce": ms_nocomm / steps,
                "exposed_comm_ms_per_step": max(ms - ms_nocomm, 0.0) / steps}
    step_eager()
    ms_eager = dx.timed(step_eager, steps)
    ops.conv_profile = []
    step_eager()
    torch.cuda.synchronize()
    prof, ops.conv_profile = ops.conv_profile, None
    conv_ms = sum(a.elapsed_time(b) for a, b, _, _ in prof)
    conv_flops = sum(f for _, _, f, _ in prof)
    by_kind = {k: {"launches": v[0], "ms": v[1], "tflops": v[2] / (v[1] * 1e-3) / 1e12 if v[1] > 0 else 0.0}
               for k, v in summarize_profile(prof, lambda key: str(key[0])).items()}
    peak_tf, _, peak_kind = measured_peaks()
    achieved = conv_flops / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
    step_flops = 5047.7e9  # fwd + dgrad + wgrad of one 128^3 patch
    block = {"metric": "train patches/s (EquiUNet-ASPP-Evo, 128^3, batch 1/GPU)", "value": world * steps / (ms * 1e-3),
             "unit": "patches/s", "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": ms / steps,
             "scaling": "weak", "dtype": "bf16", "loss_after": loss_value,
             "api": "brats21_b200.engine.TrainStep (engine.train_step replayed from one CUDA graph)",
             "ms_per_step_eager": ms_eager / steps,
             "e2e": {"value": world * steps / (ms_e2e * 1e-3), "unit": "patches/s", "ms_per_step": ms_e2e / steps,
                     "h2d_bytes_per_step": (host_img.numel() + host_tgt.numel()) * 4, "d2h_bytes_per_step": 4},
             "gpu_launches": launches, "comm": comm,
             "step_tflops": step_flops / (ms / steps * 1e-3) / 1e12,
             "step_frac_of_peak": step_flops / (ms / steps * 1e-3) / 1e12 / peak_tf,
             "roofline": {"bound": "tensor", "kernel": "conv3d fwd + dgrad + wgrad (tcgen05), all launches of one step",
                          "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
                          "peak_kind": f"{peak_kind} bf16 (sustained)", "traffic": None, "launches": len(prof),
                          "by_kernel": by_kind,
                          "flops_per_launch": conv_flops / max(len(prof), 1), "avg_launch_ms": conv_ms / max(len(prof), 1),
                          "share_of_step": conv_ms / (ms / steps) if ms > 0 else None}}
    del net, model, opt
    torch.cuda.empty_cache()
    return block


def run_b200_train(args, wl, dx: Dist):
    sampler = ClockSampler(dx.local_rank)
    sampler.start()
    block = measure_train(args, dx, args.steps, args.warmup, wl["seed"])
    clocks = sampler.stop()
    line = {k: block[k] for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step")}
    line.update({"higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                 "config": workload_config(args, wl), "clocks": clocks, "e2e": block["e2e"],
                 "gpu_launches": block["gpu_launches"], "roofline": block["roofline"], "comm": block["comm"],
                 "step_tflops": block["step_tflops"], "step_frac_of_peak": block["step_frac_of_peak"]})
    if dx.rank == 0:
        print(json.dumps(line), flush=True)


def load_profile_json(name):
    p = os.path.join(ROOT, "profiles", name)
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f)
    return None


def build_models(wl, dev):
    import warnings
    from brats21_b200 import networks
    feats = [WIDTH * 2 ** i for i in range(4)]
    cls = networks.EquiUnetASSPEvo if wl["version"] == 2 else networks.EquiUnet
    nets = []
    for seed in wl.get("ensemble", (wl["seed"],)):  # the reference's seeds (arguments_train.py:103) + 7
        torch.manual_seed(seed)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            nets.append(cls(4, 3, feats, norm_layer="group", act="relu", deep_supervision=True).to(dev).eval())
    return nets


def run_b200_cohort(args, wl, dx: Dist):
    """BASELINE configs[4]: the whole validation cohort, volumes sharded round-robin over the ranks (no collective on the
    data path).  One step = the whole cohort; a host thread generates the next synthetic volume while the GPU works (the
    reference's DataLoader workers), H2D of the volume and D2H of the labels are inside the timed region."""
    import queue
    from brats21_b200 import _lib, engine, parallel, synth, tta
    dev = dx.dev
    nets = build_models(wl, dev)
    comp = tta.get_flip8_transforms()
    mine = parallel.shard_indices(args.cohort, dx.rank, dx.world)
    host_lab = torch.empty((1, 1) + VOL_SHAPE, dtype=torch.uint8).pin_memory()
    pinned = [torch.empty((1, 4) + VOL_SHAPE, dtype=torch.float32).pin_memory() for _ in range(3)]

    def producer(q):
        for j, idx in enumerate(mine):
            buf = pinned[j % 3]
            buf.copy_(synth.volume(seed=idx, shape=VOL_SHAPE))
            q.put(buf)
        q.put(None)

    def one(buf):
        v, meta = pad_to_8(buf.to(dev, non_blocking=True))
        _, label = engine.predict_volume(nets, v, comp, True, ROI, wl["sw_batch"], 0.25, wl["mode"])
        crop = label[..., meta[0][0]:meta[0][0] + meta[0][1], meta[1][0]:meta[1][0] + meta[1][1],
                     meta[2][0]:meta[2][0] + meta[2][1]]
        host_lab.copy_(crop, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return int(host_lab.sum())  # the consumer of the labels (keeps the D2H honest)

    warm = synth.volume(seed=10_000 + dx.rank, shape=VOL_SHAPE).pin_memory()
    one(warm)
    sampler = ClockSampler(dx.local_rank)
    q = queue.Queue(maxsize=2)
    th = threading.Thread(target=producer, args=(q,), daemon=True)
    th.start()
    first = q.get()  # the first volume is ready before the clock starts (a DataLoader would have prefetched it)
    dx.barrier()
    sampler.start()
    l0 = _lib.launch_count
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    buf, done, checksum = first, 0, 0
    while buf is not None:
        checksum += one(buf)
        done += 1
        buf = q.get()
    e1.record()
    dx.barrier()
    wall = dx.max_over_ranks(time.perf_counter() - t0)
    ms = dx.max_over_ranks(e0.elapsed_time(e1))
    launches = _lib.launch_count - l0
    clocks = sampler.stop()
    value = args.cohort / (ms * 1e-3)
    line = {"metric": "volumes/s (sliding-window + 8xTTA)", "value": value, "unit": "volumes/s", "n_gpus": dx.world,
            "steps": 1, "warmup": 1, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": workload_config(args, wl),
            "clocks": clocks, "cohort": {"volumes": args.cohort, "volumes_this_rank": done, "device_s": ms * 1e-3,
                                         "wall_s": wall, "windows": args.cohort * windows_per_volume(wl)},
            "e2e": {"value": args.cohort / wall, "unit": "volumes/s", "ms_per_step": wall * 1e3,
                    "h2d_bytes_per_step": args.cohort * pinned[0].numel() * 4,
                    "d2h_bytes_per_step": args.cohort * host_lab.numel()},
            "gpu_launches": launches}
    if dx.rank == 0:
        print(json.dumps(line), flush=True)


def run_b200(args, wl):
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (no CPU fallback for the product path)")
    dx = Dist()
    try:
        if wl.get("train"):
            return run_b200_train(args, wl, dx)
        if wl.get("cohort"):
            return run_b200_cohort(args, wl, dx)
        return run_b200_infer(args, wl, dx)
    finally:
        dx.close()


def run_b200_infer(args, wl, dx: Dist):
    from brats21_b200 import _lib, engine, ops, synth, tta
    dev, rank, world = dx.dev, dx.rank, dx.world
    nets = build_models(wl, dev)
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

    for _ in range(args.warmup):
        step_device()
    sampler = ClockSampler(dx.local_rank)
    sampler.start()
    l0 = _lib.launch_count
    ms = dx.timed(step_device, args.steps)
    launches = _lib.launch_count - l0
    clocks = sampler.stop()
    step_e2e()
    ms_e2e = dx.timed(step_e2e, args.steps)
    # one more step, launched eagerly with a CUDA-event pair around every conv / HBM-bound launch (the timed steps
    # above replay the per-batch forward from a CUDA graph): live per-launch durations for the two rooflines
    ops.conv_profile, ops.hbm_profile = [], []
    step_device()
    torch.cuda.synchronize()
    prof, ops.conv_profile = ops.conv_profile, None
    hprof, ops.hbm_profile = ops.hbm_profile, None

    # dominant kernel family (conv implicit GEMM): algorithmic FLOPs / event-timed launch durations
    conv_ms = sum(a.elapsed_time(b) for a, b, _, _ in prof)
    conv_flops = sum(f for _, _, f, _ in prof)
    peak_tf, peak_gbs, peak_kind = measured_peaks()
    achieved = conv_flops / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
    by_kind = {k: {"launches": v[0], "ms": v[1], "tflops": v[2] / (v[1] * 1e-3) / 1e12 if v[1] > 0 else 0.0}
               for k, v in summarize_profile(prof, lambda key: str(key[0])).items()}
    # DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture of THIS round's build
    traffic = load_profile_json("r02_ncu_traffic.json") or {}
    tk = traffic.get("dominant", {})
    roofline = {"bound": "tensor", "kernel": "conv3d implicit GEMM (tcgen05): all conv launches of one step "
                "(march = plane-marching 48-ch layers, slide = sliding-window 96-ch layers, tap = generic, point = 1x1)",
                "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
                "peak_kind": f"{peak_kind} bf16 (sustained)", "by_kernel": by_kind,
                "traffic": tk.get("dram_bytes_per_launch"), "traffic_kernel": tk.get("kernel"),
                "traffic_algorithmic": tk.get("algorithmic_bytes_per_launch"), "traffic_source": tk.get("source"),
                "launches": len(prof), "flops_per_launch": conv_flops / max(len(prof), 1),
                "avg_launch_ms": conv_ms / max(len(prof), 1),
                "share_of_step": conv_ms / (ms / args.steps) if ms > 0 else None}
    hbm_ms = sum(a.elapsed_time(b) for a, b, _, _ in hprof)
    hbm_bytes = sum(f for _, _, f, _ in hprof)
    hbm = {"bound": "hbm", "peak": peak_gbs, "unit": "GB/s", "peak_kind": f"{peak_kind} copy bandwidth",
           "achieved": hbm_bytes / (hbm_ms * 1e-3) / 1e9 if hbm_ms > 0 else 0.0,
           "share_of_step": hbm_ms / (ms / args.steps) if ms > 0 else None,
           "by_kernel": {k: {"launches": v[0], "ms": v[1], "gbs": v[2] / (v[1] * 1e-3) / 1e9 if v[1] > 0 else 0.0,
                             "frac": v[2] / (v[1] * 1e-3) / 1e9 / peak_gbs if v[1] > 0 else 0.0}
                         for k, v in summarize_profile(hprof).items()}}
    hbm["frac"] = hbm["achieved"] / peak_gbs

    value = world * args.steps / (ms * 1e-3)
    e2e_value = world * args.steps / (ms_e2e * 1e-3)
    line = {"metric": "volumes/s (sliding-window + 8xTTA)", "value": value, "unit": "volumes/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": workload_config(args, wl),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "volumes/s", "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": host_vol.numel() * 4, "d2h_bytes_per_step": host_lab.numel()},
            "gpu_launches": launches, "roofline": roofline, "hbm": hbm}
    parity = load_profile_json("r02_parity_full_size.json")
    if parity is not None:
        line["parity"] = {"source": "profiles/r02_parity_full_size.json (tests/test_gpu_fullsize.py on a B200: CUDA path "
                                    "vs the fp32 oracle on identical synthetic inputs, width 48, full sizes)", **parity}
    if not args.no_train and args.workload == "v2_tta8":
        del nets
        torch.cuda.empty_cache()
        line["train"] = measure_train(args, dx, max(10 * args.steps, 20), max(args.warmup, 3) + 2)
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            nwin = windows_per_volume(wl)
            (t,), kind = cpu_window_forward_seconds(wl["version"], 1)
            line["cpu_baseline"] = {"value": 1.0 / (nwin * t), "unit": "volumes/s", "cores": os.cpu_count(),
                                    "kind": kind, "sample": f"1 of {nwin} windows (V{wl['version']} forward 128^3 "
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="v2_tta8", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the `train` block of the default workload")
    ap.add_argument("--cohort", type=int, default=219, help="volumes of the v2_ens3_tta8_cohort workload")
    ap.add_argument("--sw-batch", type=int, default=0, help="override the workload's sliding-window batch size")
    args = ap.parse_args()
    rank, local_rank, world = dist_env()
    wl = dict(WORKLOADS[args.workload])
    if args.sw_batch > 0 and "sw_batch" in wl:
        wl["sw_batch"] = args.sw_batch
    if args.impl == "reference":
        run_reference(args, wl, rank, world)
    else:
        run_b200(args, wl)


if __name__ == "__main__":
    main()
