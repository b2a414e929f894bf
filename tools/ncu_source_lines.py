#!/usr/bin/env python
"""Per-source-line summary of one `ncu --set full --import-source on` capture (needs -lineinfo; runs without a GPU):
warp-stall samples, executed warp instructions and the leading stall reasons of the hottest lines.

    python tools/ncu_source_lines.py gpurun_out/r02h_decode_select_kernel.ncu-rep [top_n] > profiles/<name>_lines.txt
"""
import collections, csv, io, subprocess, sys

rep = sys.argv[1]
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
cur, hdr, out, kernel = None, None, [], ""
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur, hdr = r[1].split("/")[-1], None
    elif r[0] == "Function Name":
        kernel = r[1].split("(")[0]
    elif r[0] == "Line No":
        hdr = r
    elif hdr and cur and r[0].isdigit():
        d = dict(zip(hdr[4:], r[4:]))          # (the first "Source" column is the CUDA text, the second the SASS text)
        try:
            smp, ins = int(d["# Samples"]), int(d["Instructions Executed"])
        except (KeyError, ValueError):
            continue
        stalls = {k[6:]: int(v) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and v.isdigit() and int(v)}
        out.append((cur, int(r[0]), r[1].strip(), smp, ins, stalls))
tot_s, tot_i = sum(o[3] for o in out) or 1, sum(o[4] for o in out) or 1
print(f"# {rep.split('/')[-1]}: {kernel}")
print(f"# {tot_s} warp-stall samples, {tot_i} executed warp instructions; share of samples / of instructions per source line")
agg = collections.Counter()
for o in out:
    for k, v in o[5].items():
        agg[k] += v
print("# stall reasons over the whole kernel: " + ", ".join(f"{k} {100 * v / sum(agg.values()):.1f} %" for k, v in agg.most_common(8)))
per_file = collections.Counter()
for o in out:
    per_file[o[0]] += o[3]
print("# by file: " + ", ".join(f"{k} {100 * v / tot_s:.1f} %" for k, v in per_file.most_common()))
for o in sorted(out, key=lambda o: -o[3])[:top_n]:
    top = ", ".join(f"{k} {v}" for k, v in sorted(o[5].items(), key=lambda kv: -kv[1])[:3])
    print(f"{o[0]}:{o[1]:<5d} {100 * o[3] / tot_s:5.1f} % smp {100 * o[4] / tot_i:5.1f} % ins  [{top}]  {o[2][:100]}")
