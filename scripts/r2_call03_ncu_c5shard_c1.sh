#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== ncu: warp-stream kernel on the scale-27 shard (x = 1.07 GB)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmv_warp_stream -s 3 -c 1 \
  -o gpurun_out/r2_prof_c5shard -f python scripts/exp_r2.py spmv c5shard 3 > gpurun_out/r2_prof_c5shard.log 2>&1
tail -2 gpurun_out/r2_prof_c5shard.log
echo "== ncu: C1 with and without the evict_first hint on A"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:spmv_warp_stream -s 8 -c 1 \
  -o gpurun_out/r2_prof_c1_base -f python scripts/exp_r2.py spmv c1 8 > gpurun_out/r2_prof_c1_base.log 2>&1
SPBLAS_B200_LIB=$PWD/spblas_reference_b200/libspblas_b200_ef.so timeout 300 ncu --set full --clock-control none --import-source on -k regex:spmv_warp_stream -s 8 -c 1 \
  -o gpurun_out/r2_prof_c1_ef -f python scripts/exp_r2.py spmv c1 8 > gpurun_out/r2_prof_c1_ef.log 2>&1
tail -2 gpurun_out/r2_prof_c1_ef.log
ls -la gpurun_out/*.ncu-rep
