"""Turn the ncu outputs of one bench.py command into the committed summaries under profiles/.
Usage: python tools/summarize_ncu.py <cells> <launches.csv> <full.ncu-rep> [more.ncu-rep ...]"""
import collections
import csv
import subprocess
import sys

cells, launches, reps = sys.argv[1], sys.argv[2], sys.argv[3:]
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
with open(f"profiles/r1_launch_list_summary_{cells}.csv", "w") as f:
    f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none over `python bench.py --steps 2 --warmup 3 --no-cpu-baseline "
            f"--no-e2e` ({cells} cells per GPU); cold-cache, serialised: compare SHARES with bench.py's live shares\n")
    f.write(f"# launches captured: {sum(v[0] for v in agg.values())}\nkernel,launches,total_ms,share\n")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"\"{k}\",{n},{t:.3f},{t / tot:.4f}\n")

want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__grid_size", "launch__block_size"]
with open(f"profiles/r1_top_kernels_ncu_full_{cells}.csv", "w") as f:
    w = csv.writer(f)
    w.writerow(["kernel", "duration_ms", "dram_read_bytes", "dram_write_bytes", "dram_throughput_pct_of_peak", "registers_per_thread",
                "warps_active_pct", "grid", "block"])
    for rep in reps:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        hdr, units = rows[0], rows[1]
        col = {n: i for i, n in enumerate(hdr)}
        for r in rows[2:]:
            vals = []
            for m in want:
                v = float(r[col[m]].replace(",", ""))
                vals.append(v * UNIT.get(units[col[m]], 1.0))
            w.writerow([r[col["Kernel Name"]].split("(")[0].replace("void ", "")] + [f"{v:.6g}" for v in vals])
print(open(f"profiles/r1_top_kernels_ncu_full_{cells}.csv").read())
