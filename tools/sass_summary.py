#!/usr/bin/env python
"""Opcode summary of the shipped library's SASS (cuobjdump -sass): per kernel, the instruction count and the opcodes that show what
the kernel is made of -- FP64 arithmetic (DFMA/DADD/DMUL/MUFU.RCP64H), bulk asynchronous copies and their barriers (UBLKCP, SYNCS),
warp exchange (SHFL), atomics and fences.  No tensor-core opcodes are expected: nothing on this path is a dense contraction.

    python tools/sass_summary.py > profiles/r2_sass_opcodes.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
so = os.path.join(ROOT, "additivefoam_b200", "libadditivefoam_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
dem = {}
kern = None
ops = collections.OrderedDict()
arch = set()
for line in out.splitlines():
    m = re.match(r"\s*arch = (\S+)", line)
    if m:
        arch.add(m.group(1))
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = m.group(1)
        ops[kern] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and kern:
        ops[kern][m.group(1)] += 1
names = subprocess.run(["c++filt"], input="\n".join(ops), capture_output=True, text=True).stdout.splitlines()
KEY = ("DFMA", "DADD", "DMUL", "MUFU.RCP64H", "DSETP", "UBLKCP", "SYNCS", "SHFL", "ATOM", "RED", "MEMBAR", "FENCE", "LDG", "STG", "LDS", "STS",
       "LD.E", "ST.E", "BAR", "HMMA", "UTCHMMA", "UTCMMA", "IMMA", "DMMA", "NANOSLEEP", "BRA")
print(f"# cuobjdump -sass {os.path.relpath(so, ROOT)}: arch {', '.join(sorted(arch))}; {len(ops)} kernels")
print("# kernel | instructions | key opcodes (prefix match: count)")
tot = collections.Counter()
for k, n in zip(ops, names):
    c = ops[k]
    total = sum(c.values())
    short = re.sub(r"\(.*", "", n).replace("void ", "").replace("b200::", "")
    agg = collections.Counter()
    for op, v in c.items():
        for key in KEY:
            if op.startswith(key):
                agg[key] += v
                tot[key] += v
                break
    print(f"{short} | {total} | " + ", ".join(f"{a}:{b}" for a, b in agg.most_common() if a != "BRA"))
print("# whole library: " + ", ".join(f"{a}:{b}" for a, b in tot.most_common()))
print("# tensor-core opcodes (HMMA/IMMA/DMMA/UTC*MMA): " + str(sum(tot[k] for k in ("HMMA", "IMMA", "DMMA", "UTCHMMA", "UTCMMA"))))
