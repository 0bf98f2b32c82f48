#!/bin/bash
# DRAM bytes of the SpMM kernel on C5 (scale 27, k = 32, fp64, one GPU): one-pass metrics only —
# a full-set replay would save and restore the 34 GB C around every pass
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
SPBLAS_B200_BENCH_VERBOSE=1 timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct \
  --clock-control none -k regex:'spmm_ring|spmm_row' -s 2 -c 2 --csv --log-file gpurun_out/r2_ncu_c5mm_traffic.csv \
  python bench.py --configs c5 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2_ncu_c5mm_traffic.log 2>&1
tail -12 gpurun_out/r2_ncu_c5mm_traffic.csv | cut -c1-260
