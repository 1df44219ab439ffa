#!/bin/bash
# thin-slab sweep timings of the one-tile-per-CTA kernel for each ring geometry (one B200), inside a PCG/DIC solve
cd "$(dirname "$0")/.."
for ring in ${RINGS:-86 412 48 83 44}; do
  echo "== B200_WAVE=${WAVE:-single} B200_WAVE2_RING=$ring (1000x400x31)"
  B200_WAVE=${WAVE:-single} B200_WAVE2_RING=$ring timeout 120 python tools/perf_ldu.py 1000 400 31 5 solve || echo "FAILED rc=$?"
done
for g in "1000 8 4" "1000 400 4" "1000 8 31"; do
  echo "== ${WAVE:-single} default ring, $g"; B200_WAVE=${WAVE:-single} timeout 60 python tools/perf_ldu.py $g 5 solve
done
