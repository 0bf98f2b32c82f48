#!/bin/bash
# A/B: the global-hub walk with L2 eviction hints (table evict_last, A and cold gathers evict_first)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for lib in base l2keep; do
  L=""; [ $lib != base ] && L=$PWD/spblas_reference_b200/libspblas_b200_$lib.so
  SPBLAS_B200_LIB=$L EXP_MATRIX_OPT=1 timeout 400 python scripts/exp_r2.py spmv c5shard 20 2>&1 | cut -c1-300
  SPBLAS_B200_LIB=$L SPBLAS_B200_BENCH_VERBOSE=1 timeout 900 python bench.py --configs c5 --no-cpu-baseline --no-e2e > gpurun_out/r2_c5_n1_$lib.json 2> gpurun_out/r2_c5_n1_$lib.err
  grep "c5: step\|c5mm" gpurun_out/r2_c5_n1_$lib.err
done 2>&1 | tee gpurun_out/r2_hubg_l2keep.txt
