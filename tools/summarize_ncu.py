"""Summarise ncu output brought back from the GPU box into small text files under profiles/.

    python tools/summarize_ncu.py launches gpurun_out/r01_launches.csv profiles/r01_launches.md [title]
    python tools/summarize_ncu.py full gpurun_out/r01_conv.ncu-rep profiles/r01_conv_full.md [title]

`launches` aggregates a `--metrics gpu__time_duration.sum --csv` launch list per kernel (count, total, share).
`full` extracts the roofline-relevant raw metrics of every captured launch from a `--set full` report.
"""
import collections
import csv
import subprocess
import sys

FULL_KEYS = [
    "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__ops_path_tensor_op_hmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sectors_srcunit_tex.sum",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "smsp__cycles_active.avg",
]


def launches(src, dst, title):
    rows = list(csv.reader(open(src)))
    hdr, agg = None, collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        if r and r[0] == "ID":
            hdr = r
            continue
        if hdr is None or len(r) < len(hdr):
            continue
        d = dict(zip(hdr, r))
        if d["Metric Name"] != "gpu__time_duration.sum":
            continue
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(d["Metric Unit"], 1e-6)
        name = d["Kernel Name"].split("(")[0].replace("void ", "")
        agg[name][0] += 1
        agg[name][1] += float(d["Metric Value"].replace(",", "")) * scale
    tot = sum(v[1] for v in agg.values())
    with open(dst, "w") as f:
        f.write(f"# {title}\n\nsource: `{src}` (ncu --metrics gpu__time_duration.sum --clock-control none; per-launch "
                f"times are cold-cache and serialised: compare shares, not absolutes)\n\n")
        f.write(f"total {tot:.2f} ms over {sum(v[0] for v in agg.values())} launches\n\n")
        f.write("| kernel | launches | total ms | share | avg us |\n|---|---:|---:|---:|---:|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {v[0]} | {v[1]:.2f} | {v[1] / tot:.3f} | {1e3 * v[1] / v[0]:.1f} |\n")
    print(open(dst).read())


def full(src, dst, title):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write(f"# {title}\n\nsource: `{src}` (ncu --set full --clock-control none --import-source on)\n\n")
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            u = dict(zip(hdr, units))
            f.write(f"## launch {d.get('ID')}: `{d.get('Kernel Name', '').split('(')[0]}`\n\n")
            for k in FULL_KEYS:
                if k in d:
                    f.write(f"- {k} = {d[k]} {u.get(k, '')}\n")
            try:
                t = float(d["dram__bytes_read.sum"]) + float(d["dram__bytes_write.sum"])
                f.write(f"- traffic (dram read + write) = {t:.3f} {u.get('dram__bytes_read.sum', '')}\n")
            except (KeyError, ValueError):
                pass
            f.write("\n")
    print(open(dst).read())


if __name__ == "__main__":
    mode, src, dst = sys.argv[1:4]
    title = sys.argv[4] if len(sys.argv) > 4 else src
    (launches if mode == "launches" else full)(src, dst, title)
