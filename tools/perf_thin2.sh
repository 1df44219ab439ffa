#!/bin/bash
# thin-slab sweeps: k_wave4 (tri) against k_wave3 (multi: NW tiles per CTA) geometries, inside a PCG/DIC solve
cd "$(dirname "$0")/.."
run() { echo "== $1 $2 $3 ($4)"; B200_WAVE=$1 B200_WAVE2_RING=$2 B200_WAVE3=$3 timeout 120 python tools/perf_ldu.py $4 5 solve | grep -E "sweep|factor" || echo "FAILED rc=$?"; }
for g in "1000 400 31" "1000 8 4" "1000 400 4"; do
  run tri 86 x "$g"
  for c3 in 446 444 248 146; do run multi x $c3 "$g"; done
done
run tri 83 x "1000 400 250"
run single 83 x "1000 400 250"
