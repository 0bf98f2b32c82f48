#!/bin/bash
# ncu --set full of the warp-stream kernel (branch-free load phase) on C1
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:spmv_warp_stream -s 8 -c 1 \
  -o gpurun_out/r2_prof_c1_flat -f python scripts/exp_r2.py spmv c1 8 > gpurun_out/r2_prof_c1_flat.log 2>&1
tail -2 gpurun_out/r2_prof_c1_flat.log
