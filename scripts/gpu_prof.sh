#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmv_merge_tile -s 4 -c 2 -o gpurun_out/prof_c2_v1 -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"spmv|spmm|merge|rowlen|segments|partition" -c 40 --csv --log-file gpurun_out/launches_c2.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_bench.log 2>&1
tail -3 gpurun_out/ncu_full.log
