"""Stream timeline of one eager v2_train step from CUPTI (torch.profiler chrome trace): per stream busy time, overlap of
the main and the weight-gradient stream, idle gaps, and the ordered kernel list with start offsets.
    python tools/timeline.py [out.md]"""
import json
import os
import sys
import tempfile
import warnings

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from brats21_b200 import engine, networks, synth  # noqa: E402
from brats21_b200.losses import DiceLoss  # noqa: E402
from brats21_b200.optimizer import Ranger2020  # noqa: E402


def union(iv):
    iv = sorted(iv)
    out, cur = [], None
    for a, b in iv:
        if cur is None or a > cur[1]:
            if cur:
                out.append(cur)
            cur = [a, b]
        else:
            cur[1] = max(cur[1], b)
    if cur:
        out.append(cur)
    return out


def inter(a, b):
    i = j = 0
    tot = 0.0
    while i < len(a) and j < len(b):
        lo, hi = max(a[i][0], b[j][0]), min(a[i][1], b[j][1])
        if hi > lo:
            tot += hi - lo
        if a[i][1] < b[j][1]:
            i += 1
        else:
            j += 1
    return tot


def main():
    dev = torch.device("cuda:0")
    torch.manual_seed(93)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        net = networks.EquiUnetASSPEvo(4, 3, [48, 96, 192, 384], norm_layer="group", act="relu", deep_supervision=True).to(dev)
    net.train()
    opt = Ranger2020([p for n, p in net.named_parameters() if not n.endswith(".v")], lr=3e-4, weight_decay=1e-5, use_gc=False)
    crit = DiceLoss()
    img = synth.volume(seed=2000, shape=(128, 128, 128)).to(dev)
    tgt = synth.target(shape=(128, 128, 128)).to(dev)
    graphed = engine.TrainStep(net, crit, opt)
    for _ in range(5):
        graphed(img, tgt)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        graphed(img, tgt)
        torch.cuda.synchronize()
    path = os.path.join(tempfile.mkdtemp(), "trace.json")
    prof.export_chrome_trace(path)
    ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memset", "gpu_memcpy")]
    ev.sort(key=lambda e: e["ts"])
    t0 = ev[0]["ts"]
    span = max(e["ts"] + e["dur"] for e in ev) - t0
    streams = {}
    for e in ev:
        streams.setdefault(e["args"].get("stream"), []).append((e["ts"] - t0, e["ts"] - t0 + e["dur"]))
    out = [f"# v2_train (CUDA-graph replay): GPU span {span / 1e3:.2f} ms, {len(ev)} launches\n"]
    out.append("| stream | launches | busy ms | share of span |\n|---|---:|---:|---:|")
    un = {}
    for s, iv in sorted(streams.items(), key=lambda kv: -len(kv[1])):
        un[s] = union(iv)
        busy = sum(b - a for a, b in un[s])
        out.append(f"| {s} | {len(iv)} | {busy / 1e3:.2f} | {busy / span:.2f} |")
    allu = union([x for iv in streams.values() for x in iv])
    out.append(f"\nany stream busy: {sum(b - a for a, b in allu) / 1e3:.2f} ms; idle: {(span - sum(b - a for a, b in allu)) / 1e3:.2f} ms")
    keys = sorted(streams, key=lambda s: -len(streams[s]))
    if len(keys) > 1:
        out.append(f"main and side stream both busy: {inter(un[keys[0]], un[keys[1]]) / 1e3:.2f} ms")
    out.append("\n| start us | dur us | stream | kernel |\n|---:|---:|---|---|")
    for e in ev:
        out.append(f"| {e['ts'] - t0:.0f} | {e['dur']:.0f} | {e['args'].get('stream')} | `{e['name'].split('(')[0].replace('void ', '')[:70]}` |")
    text = "\n".join(out)
    if len(sys.argv) > 1:
        open(sys.argv[1], "w").write(text)
    print("\n".join(out[:12]))


if __name__ == "__main__":
    main()
