"""Turn the ncu outputs of one bench.py command into the committed summaries under profiles/ and into profiles/r2_traffic.json,
the per-launch DRAM bytes bench.py reports as `roofline.traffic` -- keyed by the hash of the CUDA sources, so that a number taken on
another build is never reported.
Usage: python tools/summarize_ncu.py <cells per GPU> <launches.csv> <full.ncu-rep> [more.ncu-rep ...]     (prefix: env PREFIX, default r2)"""
import collections
import csv
import re
import subprocess
import sys

import json
import os

cells, launches, reps = sys.argv[1], sys.argv[2], sys.argv[3:]
PREFIX = os.environ.get("PREFIX", "r2")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}

rows = [r for r in csv.reader(open(launches)) if r and not r[0].startswith("==")]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    try:
        v = float(r[vi].replace(",", "")) * UNIT[r[ui]]
    except Exception:
        continue
    name = r[ki].split("(")[0].replace("void ", "")
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
with open(f"profiles/{PREFIX}_launch_list_summary_{cells}.csv", "w") as f:
    f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none over `python bench.py --steps 2 --warmup 3 --no-cpu-baseline "
            f"--no-e2e` ({cells} cells per GPU); cold-cache, serialised: compare SHARES with bench.py's live shares\n")
    f.write(f"# launches captured: {sum(v[0] for v in agg.values())}\nkernel,launches,total_ms,share\n")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"\"{k}\",{n},{t:.3f},{t / tot:.4f}\n")

want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__grid_size", "launch__block_size"]
traffic_rows = []
with open(f"profiles/{PREFIX}_top_kernels_ncu_full_{cells}.csv", "w") as f:
    w = csv.writer(f)
    w.writerow(["kernel", "duration_ms", "dram_read_bytes", "dram_write_bytes", "dram_throughput_pct_of_peak", "registers_per_thread",
                "warps_active_pct", "grid", "block"])
    for rep in reps:
        # a capture exported on the GPU box (`ncu -i x.ncu-rep --page raw --csv > x_raw.csv`), or the .ncu-rep itself
        out = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        hdr, units = rows[0], rows[1]
        col = {n: i for i, n in enumerate(hdr)}
        for r in rows[2:]:
            vals = []
            for m in want:
                v = float(r[col[m]].replace(",", ""))
                vals.append(v * UNIT.get(units[col[m]], 1.0))
            kname = r[col["Kernel Name"]].split("(")[0].replace("void ", "")
            w.writerow([kname] + [f"{v:.6g}" for v in vals])
            traffic_rows.append((kname, vals[1] + vals[2]))
print(open(f"profiles/{PREFIX}_top_kernels_ncu_full_{cells}.csv").read())

# ---- per-launch DRAM traffic of the kernels bench.py's roofline can name, for THIS build
def group_of(k):
    m = re.search(r"k_wave[234]<(?:\(int\))?([012])\b", k)        # first template argument: 0 factor, 1 forward, 2 backward sweep
    if m:
        return ("factor", "sweep_fwd", "sweep_bwd")[int(m.group(1))]
    for pat, g in (("k_amul_block", "amul"), ("k_corrector", "corrector"), ("k_assemble_implicit", "assemble"), ("k_props", "props"),
                   ("k_dinum", "dinum"), ("k_faces", "faces"), ("k_alpha_update", "alpha"), ("k_normfactor_block", "normfactor")):
        if pat in k:
            return g
    return None


import bench  # noqa: E402  (source_hash)
per = collections.defaultdict(list)
for k, b in traffic_rows:
    g = group_of(k)
    if g and b > 1e6:                       # the instantiations that return at once (exact factor sweep) move nothing
        per[g].append(b)
path = os.path.join(ROOT, "profiles", f"{PREFIX}_traffic.json")
data = {"source_hash": bench.source_hash(), "bytes_per_launch": {}}
if os.path.exists(path):
    try:
        old = json.load(open(path))
        if old.get("source_hash") == data["source_hash"]:
            data = old
    except Exception:
        pass
data["bytes_per_launch"][str(cells)] = {g: sum(v) / len(v) for g, v in per.items()}
json.dump(data, open(path, "w"), indent=1)
print(open(path).read())
