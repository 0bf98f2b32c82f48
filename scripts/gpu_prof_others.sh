#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for w in c1 c4; do
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"spmv_merge_tile|spmv_pipe" -s 2 -c 1 -o gpurun_out/prof_$w -f python bench.py --workload $w --steps 3 --warmup 3 > gpurun_out/ncu_$w.log 2>&1
done
for w in c3k32 c3k128; do
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"spmm_row" -s 3 -c 1 -o gpurun_out/prof_$w -f python bench.py --workload $w --steps 3 --warmup 3 > gpurun_out/ncu_$w.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
