#!/bin/bash
# Multi-GPU evidence run on one box: bench.py on N GPUs (strong scaling of the 100 M-cell case + weak sub-record + parity_check),
# the decomposed parity run against the oracle's N-rank emulation, and (N = 8) BASELINE configs[4].  Usage: tools/run_scale.sh N TAG
cd "$(dirname "$0")/.."
N=${1:-2}; TAG=${2:-r2}; P=29700
tr() { port=$((P++)); timeout ${TMO:-500} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port "$@"; }
mkdir -p gpurun_out
tr bench.py --gpus $N --steps 8 --warmup 3 > gpurun_out/${TAG}_bench$N.json 2> gpurun_out/${TAG}_bench$N.err
TMO=300 tr tools/run_multi.py implicit 30 > gpurun_out/${TAG}_multi$N.log 2>&1
if [ "$N" = "8" ]; then
  tr bench.py --gpus 8 --steps 6 --warmup 3 --workload pbf --nx 1000 --ny 500 --nz 400 --no-weak --no-e2e --no-parity-check > gpurun_out/${TAG}_pbf8.json 2> gpurun_out/${TAG}_pbf8.err
fi
tail -2 gpurun_out/${TAG}_multi$N.log; tail -c 300 gpurun_out/${TAG}_bench$N.err
