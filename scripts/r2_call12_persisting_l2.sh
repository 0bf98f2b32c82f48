#!/bin/bash
# does L2::evict_last need the persisting-L2 set-aside (cudaLimitPersistingL2CacheSize)?
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L=$PWD/spblas_reference_b200/libspblas_b200_l2keep.so
for mb in 0 32 64 79 96; do
  SPBLAS_B200_LIB=$L EXP_PERSIST_L2_MB=$mb EXP_MATRIX_OPT=1 timeout 400 python scripts/exp_r2.py spmv c5shard 20 2>&1 | cut -c1-300
done 2>&1 | tee gpurun_out/r2_hubg_persisting_l2.txt
EXP_PERSIST_L2_MB=79 EXP_MATRIX_OPT=1 timeout 400 python scripts/exp_r2.py spmv c5shard 20 2>&1 | cut -c1-300 | tee -a gpurun_out/r2_hubg_persisting_l2.txt
echo "== launch lists (our kernels only)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'spmv_|offsets_fingerprint|hub_fill' -s 30 -c 60 --csv \
  --log-file gpurun_out/r2_launches_c2.csv python bench.py --configs none --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2_launches_c2.log 2>&1
tail -4 gpurun_out/r2_launches_c2.csv | cut -c1-160
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'spmv_|offsets_fingerprint|hub_fill' -s 20 -c 60 --csv \
  --log-file gpurun_out/r2_launches_c1_noinfo.csv python bench.py --configs c1 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2_launches_c1.log 2>&1
tail -12 gpurun_out/r2_launches_c1_noinfo.csv | cut -c1-160
