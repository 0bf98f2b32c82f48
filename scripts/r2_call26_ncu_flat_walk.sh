#!/bin/bash
# ncu --set full of the shipped walk kernels after the branch-free load phase: C4 through
# matrix_opt (spmv_hub_stream_kernel), the scale-27 shard (spmv_hubg_stream_kernel + hub_fill)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
EXP_MATRIX_OPT=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:spmv_hub_stream -s 3 -c 1 \
  -o gpurun_out/r2_prof_c4_hub_flat -f python scripts/exp_r2.py spmv c4 3 > gpurun_out/r2_prof_c4_hub_flat.log 2>&1
tail -1 gpurun_out/r2_prof_c4_hub_flat.log
EXP_MATRIX_OPT=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:'spmv_hubg_stream|hub_fill' -s 6 -c 2 \
  -o gpurun_out/r2_prof_c5shard_hubg_flat -f python scripts/exp_r2.py spmv c5shard 3 > gpurun_out/r2_prof_c5shard_hubg_flat.log 2>&1
tail -1 gpurun_out/r2_prof_c5shard_hubg_flat.log
