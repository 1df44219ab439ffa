"""Summarise an ncu source page: top stalled SASS instructions per kernel.  Usage: ncu_hot.py report.ncu-rep [kernel-substr] [topN]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
sub = sys.argv[2] if len(sys.argv) > 2 else ""
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
kern, hdr, data = None, None, {}
for r in rows:
    if r and r[0] == "Kernel Name":
        kern = r[1]; data[kern] = []; hdr = None
    elif r and r[0] == "Address":
        hdr = r
    elif kern and hdr and len(r) == len(hdr):
        data[kern].append(dict(zip(hdr, r)))
for k, ins in data.items():
    if sub not in k:
        continue
    tot = sum(int(i["# Samples"] or 0) for i in ins) or 1
    print("=" * 100); print(k, "samples", tot, "instructions", len(ins))
    stall_cols = [c for c in ins[0] if c.startswith("stall_")]
    agg = {c: sum(int(i[c] or 0) for i in ins) for c in stall_cols}
    print("stall totals:", {c: v for c, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
    for n, i in sorted(enumerate(ins), key=lambda t: -int(t[1]["# Samples"] or 0))[:top]:
        st = {c[6:]: int(i[c]) for c in stall_cols if int(i[c] or 0) > 0}
        st = dict(sorted(st.items(), key=lambda kv: -kv[1])[:3])
        print(f"{n:5d} {100 * int(i['# Samples']) / tot:5.1f}%  exec={i['Instructions Executed']:>9s}  {i['Source'].strip()[:70]:70s} {st}")
