#!/bin/bash
# Sweep of the pipelined SpMV kernel shape on C2 (consumer warps, CTAs/SM, tile items, stages)
# through the tuning environment variables read at plan creation (csrc/cabi.cu).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
run() { name=$1; shift
  env "$@" timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-e2e 2>gpurun_out/sweep_$name.err | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('$name', 'ms=%.4f'%d['ms_per_step'], 'kern_ms=%.4f'%d['roofline']['kernel_ms'], 'GB/s=%.0f'%d['roofline']['achieved'], 'frac=%.3f'%d['roofline']['frac'])
" | tee -a gpurun_out/sweep.txt
}
: > gpurun_out/sweep.txt
run default
for cfg in "8 3 2048 3" "8 3 1536 3" "8 3 2048 2" "8 2 2048 4" "8 2 3072 3" "8 2 4096 2" "8 4 1536 2" "8 1 4096 4" "16 1 3072 4" "16 1 4096 4" "16 1 2048 6"; do
  set -- $cfg
  run w$1_c$2_t$3_s$4 SPBLAS_B200_CONSUMER_WARPS=$1 SPBLAS_B200_CTAS_PER_SM=$2 SPBLAS_B200_TILE_ITEMS=$3 SPBLAS_B200_STAGES=$4
done
