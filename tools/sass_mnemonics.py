#!/usr/bin/env python
"""Per-kernel SASS summary of libcbsim.so (runs without a GPU): instruction count and the mnemonics that prove the hardware paths.

    python tools/sass_mnemonics.py > profiles/<round>_sass_mnemonics.txt
"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "c-cyberbattlesim_b200", "libcbsim.so")
KEYS = ["UTCHMMA", "UTCQMMA", "UTMALDG", "LDTM", "UTCBAR", "UBLKCP", "SYNCS", "LDGSTS", "CREDUX", "REDUX", "ACQBULK", "PREEXIT", "LDG", "BAR.SYNC",
        "SHFL", "STG", "MUFU.RSQ", "ATOMS", "ATOMG", "RED", "DFMA", "HFMA2", "F2FP"]
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
print("# cuobjdump -sass c-cyberbattlesim_b200/libcbsim.so (sm_100a), per kernel: total SASS instructions and the mnemonics that prove the")
print("# hardware paths (UTCHMMA = tcgen05.mma, UTMALDG = TMA tensor load, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk,")
print("# SYNCS = mbarrier, LDGSTS = cp.async, CREDUX / REDUX = warp reduction (the row scan's maximum), ACQBULK / PREEXIT = griddepcontrol.wait / launch_dependents,")
print("# DFMA = float64 re-score)")
blocks = sass.split("Function : ")[1:]
for blk, name in zip(blocks, names):
    body = blk.split("\n", 1)[1]
    ops = collections.Counter()
    n = 0
    for line in body.split("\n"):
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if not m:
            continue
        n += 1
        op = m.group(1)
        for k in KEYS:
            if op == k or op.startswith(k + "."):
                ops[k] += 1
                break
    short = re.sub(r"\(.*", "", name)
    print(short)
    print("    instructions", n, " " + "  ".join(f"{k} {v}" for k, v in ops.items()))
