#!/bin/bash
# The round's profile evidence on one B200: the default bench line, the launch list of the same command under ncu, `ncu --set full`
# captures of the kernels of a step at 100 M cells and of the thin slab (12.8 M cells).  Every ncu pass follows the same command run
# without ncu (B200_PROFILING.md).  Outputs under gpurun_out/; tools/summarize_ncu.py turns them into profiles/.
cd "$(dirname "$0")/.."
TAG=${1:-r2}
mkdir -p gpurun_out
K='regex:k_wave2|k_wave4|k_amul_block|k_wave_pcg|k_wave_from_natural|k_wave_gather|k_corrector|k_assemble_implicit|k_faces|k_dinum|k_props|k_alpha_update|k_normfactor_block|k_power'
timeout 600 python bench.py > gpurun_out/${TAG}_bench1.json 2> gpurun_out/${TAG}_bench1.err
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_plain100.json 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${TAG}_launches_100000000.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_ncu_list100.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k "$K" -s 300 -c 70 -f -o gpurun_out/${TAG}_full_100000000 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_ncu_full100.log 2>&1
K2='regex:k_props|k_dinum|k_faces|k_assemble_implicit|k_corrector|k_alpha_update|k_power|k_qdot'
timeout 900 ncu --set full --clock-control none --import-source on -k "$K2" -s 8 -c 16 -f -o gpurun_out/${TAG}_full_step_100000000 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_ncu_full100b.log 2>&1
timeout 200 python bench.py --nz 32 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_plain12.json 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${TAG}_launches_12800000.csv python bench.py --nz 32 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_ncu_list12.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "$K" -s 300 -c 100 -f -o gpurun_out/${TAG}_full_12800000 python bench.py --nz 32 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_ncu_full12.log 2>&1
ls -la gpurun_out/${TAG}_*
