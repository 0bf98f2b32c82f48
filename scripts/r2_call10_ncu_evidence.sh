#!/bin/bash
# ncu evidence of round 2 (one GPU): launch list of the headline command, and ncu --set full of
# the kernels whose DRAM traffic goes into profiles/traffic.json
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== launch list of the headline command"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 300 --csv \
  --log-file gpurun_out/r2_launches_c2.csv python bench.py --configs none --steps 20 --warmup 3 --no-cpu-baseline --no-e2e \
  > gpurun_out/r2_launches_c2.log 2>&1
tail -3 gpurun_out/r2_launches_c2.csv | cut -c1-200
echo "== ncu --set full: global-hub kernel + fill kernel on the scale-27 shard (per-GPU product at N=8)"
EXP_MATRIX_OPT=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'spmv_hubg_stream|hub_fill' -s 6 -c 2 \
  -o gpurun_out/r2_prof_c5shard_hubg -f python scripts/exp_r2.py spmv c5shard 3 > gpurun_out/r2_prof_c5shard_hubg.log 2>&1
tail -2 gpurun_out/r2_prof_c5shard_hubg.log
echo "== ncu --set full: C5 scale 27 on one GPU (SpMV hubg kernel, SpMM ring kernel)"
SPBLAS_B200_BENCH_VERBOSE=1 timeout 1500 ncu --set full --clock-control none -k regex:'spmv_hubg_stream|spmm_ring' -s 8 -c 3 \
  -o gpurun_out/r2_prof_c5_n1 -f python bench.py --configs c5 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2_prof_c5_n1.log 2>&1
tail -4 gpurun_out/r2_prof_c5_n1.log
echo "== ncu --set full: C4 hub kernel, C1 (info and no-info: the check kernel)"
EXP_MATRIX_OPT=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:spmv_hub_stream -s 3 -c 1 \
  -o gpurun_out/r2_prof_c4_hub -f python scripts/exp_r2.py spmv c4 3 > gpurun_out/r2_prof_c4_hub.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 24 --csv --log-file gpurun_out/r2_launches_c1_noinfo.csv \
  python bench.py --configs c1 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2_launches_c1.log 2>&1
tail -8 gpurun_out/r2_launches_c1_noinfo.csv | cut -c1-220
ls -la gpurun_out/*.ncu-rep
