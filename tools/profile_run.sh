#!/bin/bash
# The round's profile evidence on one B200: the default bench line, the launch list of the same command under ncu, `ncu --set full`
# captures of the kernels of a step at 100 M cells and of the thin slab (12.8 M cells).  Every ncu pass follows the same command run
# without ncu (B200_PROFILING.md).  The captures are exported to raw-page CSV on the box and the .ncu-rep files deleted (they do not
# fit the 64 MiB that come back).  Outputs under gpurun_out/; tools/summarize_ncu.py turns them into profiles/.
cd "$(dirname "$0")/.."
TAG=${1:-r2}
mkdir -p gpurun_out
T0=$(date +%s); lap() { echo "[$(( $(date +%s) - T0 )) s] $1"; }
K='regex:k_wave2|k_wave4|k_amul_block|k_wave_pcg|k_wave_from_natural|k_wave_gather|k_corrector|k_assemble_implicit|k_faces|k_dinum|k_props|k_alpha_update|k_normfactor_block|k_power'
K2='regex:k_props|k_dinum|k_faces|k_assemble_implicit|k_corrector|k_alpha_update|k_power|k_qdot'
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
full() {   # full NAME SKIP COUNT KERNELS [bench args]
  local name=$1 skip=$2 count=$3 kern=$4; shift 4
  timeout 700 ncu --set full --clock-control none -k "$kern" -s $skip -c $count -f -o /tmp/${name} $B "$@" > gpurun_out/${name}.log 2>&1
  ncu -i /tmp/${name}.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>> gpurun_out/${name}.log
  rm -f /tmp/${name}.ncu-rep
  lap "$name: $(wc -l < gpurun_out/${name}_raw.csv) lines"
}
timeout 600 python bench.py > gpurun_out/${TAG}_bench1.json 2> gpurun_out/${TAG}_bench1.err; lap bench
timeout 300 $B > gpurun_out/${TAG}_plain100.json 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${TAG}_launches_100000000.csv $B > gpurun_out/${TAG}_ncu_list100.log 2>&1; lap "launch list 100 M"
full ${TAG}_full_100000000 300 44 "$K"
full ${TAG}_full_step_100000000 8 12 "$K2"
[ "${THIN:-1}" = "0" ] && { du -sh gpurun_out | tail -1; exit 0; }      # THIN=0: the 100 M-cell part only
timeout 200 $B --nz 32 > gpurun_out/${TAG}_plain12.json 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${TAG}_launches_12800000.csv $B --nz 32 > gpurun_out/${TAG}_ncu_list12.log 2>&1; lap "launch list 12.8 M"
full ${TAG}_full_12800000 300 44 "$K" --nz 32
du -sh gpurun_out | tail -1
