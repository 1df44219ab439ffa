#!/bin/bash
# k_wave4 on the thin slab for look-ahead depths of the halo warp (B200_WAVE4_DEPTH) and ring geometries
cd "$(dirname "$0")/.."
for d in ${DEPTHS:-1 2}; do for ring in ${RINGS:-86 84 83 48 412}; do
  echo "== tri depth $d ring $ring (1000x400x31)"
  B200_WAVE=tri B200_WAVE4_DEPTH=$d B200_WAVE2_RING=$ring timeout 120 python tools/perf_ldu.py 1000 400 31 5 solve | grep -E "sweep|factor" || echo "FAILED rc=$?"
done; done
for g in "1000 8 4" "1000 400 4" "1000 8 31"; do
echo "== tri depth 1, $g"; B200_WAVE=tri B200_WAVE4_DEPTH=1 timeout 60 python tools/perf_ldu.py $g 5 solve | grep -E "sweep|factor"
done
for ring in 83 84; do
echo "== tri depth 1 ring $ring, 100 M"; B200_WAVE=tri B200_WAVE4_DEPTH=1 B200_WAVE2_RING=$ring timeout 200 python tools/perf_ldu.py 1000 400 250 3 solve | grep -E "sweep|factor"
done
