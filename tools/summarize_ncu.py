#!/usr/bin/env python
"""Summarise ncu outputs into profiles/: a launch list CSV (--metrics gpu__time_duration.sum) and/or a full
capture (.ncu-rep, read with `ncu -i ... --page raw --csv`).

    python tools/summarize_ncu.py launches gpurun_out/launches.csv
    python tools/summarize_ncu.py full gpurun_out/prof.ncu-rep
"""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"]


def launches(path):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    d = collections.defaultdict(list)
    for r in rows[hdr + 1:]:
        if len(r) > 5:
            d[r[4].split("(")[0]].append(float(r[-1]))
    tot = sum(sum(v) for v in d.values())
    print("| kernel | launches | avg us | share of step |\n|---|---|---|---|")
    for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
        print(f"| `{k}` | {len(v)} | {sum(v) / len(v) / 1e3:.1f} | {100 * sum(v) / tot:.1f} % |")
    n = max(len(v) for v in d.values())
    print(f"\nsum of per-launch averages: {sum(sum(v) / len(v) for v in d.values()) / 1e3:.1f} us per step ({n} steps captured)")


def full(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    seen = set()
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("(")[0]
        if name in seen:
            continue
        seen.add(name)
        print(f"\n### `{name}`\n\n| metric | value | unit |\n|---|---|---|")
        for k in KEYS:
            if k in idx:
                print(f"| {k} | {r[idx[k]]} | {units[idx[k]]} |")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
