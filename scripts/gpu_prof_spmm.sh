#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
WL=${1:-c3k32}; NAME=${2:-prof_spmm}; shift; shift
env "$@" timeout 900 ncu --set full --clock-control none --import-source on -k regex:"spmm_flat|spmm_row|spmm_ring" -s 3 -c 1 -o gpurun_out/$NAME -f python bench.py --workload $WL --steps 3 --warmup 3 > gpurun_out/ncu_spmm.log 2>&1
tail -2 gpurun_out/ncu_spmm.log
