#!/bin/bash
# First GPU call of the next round: everything that was written after round 1's GPU time was
# spent, in the order that fails fastest.  Every step has its own timeout and log under
# gpurun_out/; nothing here changes a default.
#   gpurun --timeout 900 -- 'bash scripts/gpu_round2_first.sh'
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
step() { echo "== $1"; }

step "1. never-run tests (persistent triangular solve, L1-bypassing gathers)"
SPBLAS_B200_RUN_UNVALIDATED=1 timeout 240 python -m pytest tests/test_gpu_zz_unvalidated.py -x -q \
  > gpurun_out/r2_unvalidated_tests.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_unvalidated_tests.log
tail -5 gpurun_out/r2_unvalidated_tests.log

step "2. hub table size x L1 bypass on C4's matrix (and the plain walk with bypass)"
timeout 120 python scripts/hub_ab.py 24 fp32 0,0g,32768,32768g,40960g,49152g \
  > gpurun_out/r2_hub_ab_s24_fp32.log 2>&1
cat gpurun_out/r2_hub_ab_s24_fp32.log
timeout 120 python scripts/hub_ab.py 24 fp64 0,0g,12288,12288g,20480g \
  > gpurun_out/r2_hub_ab_s24_fp64.log 2>&1
cat gpurun_out/r2_hub_ab_s24_fp64.log

step "3. triangular solve: graph replay vs the persistent flag-synchronised launch"
timeout 300 python bench.py --workload trsv --steps 10 --warmup 3 > gpurun_out/r2_bench_trsv.json \
  2> gpurun_out/r2_bench_trsv.err
tail -c 1500 gpurun_out/r2_bench_trsv.json

step "4. C4 through matrix_opt (hub variant) with the plain walk beside it"
timeout 300 python bench.py --workload c4 --steps 30 --warmup 5 > gpurun_out/r2_bench_c4.json \
  2> gpurun_out/r2_bench_c4.err
tail -c 2500 gpurun_out/r2_bench_c4.json

step "5. C5's SpMM half on one GPU (fp64, k = 32, scale 22)"
timeout 300 python bench.py --workload c5mm --steps 20 --warmup 3 > gpurun_out/r2_bench_c5mm_n1.json \
  2> gpurun_out/r2_bench_c5mm_n1.err
tail -c 1500 gpurun_out/r2_bench_c5mm_n1.json

step "6. the whole GPU suite and the headline"
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1
tail -3 gpurun_out/r2_pytest_gpu.log
timeout 300 python bench.py > gpurun_out/r2_bench_c2.json 2> gpurun_out/r2_bench_c2.err
tail -c 800 gpurun_out/r2_bench_c2.json
