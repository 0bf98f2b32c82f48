#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NAME=${1:-prof_c2}
shift
env "$@" timeout 900 ncu --set full --clock-control none --import-source on -k regex:"spmv_pipe|spmv_merge_tile" -s 4 -c 1 -o gpurun_out/$NAME -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
