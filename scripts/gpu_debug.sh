#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export SPBLAS_B200_SPMV_VARIANT=1
SPBLAS_B200_STAGES=2 SPBLAS_B200_CTAS_PER_SM=1 timeout 600 compute-sanitizer --tool memcheck python scripts/debug_pipe.py 512 > gpurun_out/sanitizer.log 2>&1
grep -v "^$" gpurun_out/sanitizer.log | head -60
echo ---- plain run
SPBLAS_B200_STAGES=2 SPBLAS_B200_CTAS_PER_SM=1 timeout 120 python scripts/debug_pipe.py 512 2>&1 | tail -5
timeout 120 python scripts/debug_pipe.py 2048 2>&1 | tail -5
