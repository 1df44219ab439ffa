#!/bin/bash
# thin-slab sweep timings for every sweep schedule (one B200): ms per launch of the forward / backward / factor sweeps inside a PCG/DIC solve
cd "$(dirname "$0")/.."
for cfg in "multi 446" "multi 444" "multi 248" "multi 146" "single x"; do
  set -- $cfg
  echo "== B200_WAVE=$1 B200_WAVE3=$2  (1000x400x31)"
  B200_WAVE=$1 B200_WAVE3=$2 timeout 120 python tools/perf_ldu.py 1000 400 31 5 solve || echo "FAILED rc=$?"
done
echo "== default (100 M cells: 1000x400x250)"
timeout 300 python tools/perf_ldu.py 1000 400 250 3 solve || echo "FAILED rc=$?"
for c3 in 446 183 146; do
echo "== B200_WAVE=multi B200_WAVE3=$c3 (100 M cells)"
B200_WAVE=multi B200_WAVE3=$c3 timeout 300 python tools/perf_ldu.py 1000 400 250 3 solve || echo "FAILED rc=$?"
done
