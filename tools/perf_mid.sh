#!/bin/bash
# DIC sweeps at the slab sizes between the thin one and the full block (what 4 and 2 GPUs get of the 100 M-cell case): the
# three-role kernel (tri) against the two-warp one (single), inside a PCG/DIC solve
cd "$(dirname "$0")/.."
run() { echo "== $1 ring $2 ($3)"; B200_WAVE=$1 B200_WAVE2_RING=$2 timeout 120 python tools/perf_ldu.py $3 4 solve | grep -E "sweep|factor" || echo "FAILED rc=$?"; }
for g in "1000 400 63" "1000 400 125"; do
  run tri 86 "$g"; run tri 84 "$g"; run tri 83 "$g"
  run single 83 "$g"; run single 86 "$g"
done
