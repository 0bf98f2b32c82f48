#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
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
run v1_default SPBLAS_B200_SPMV_VARIANT=1
for cw in 8 16; do for st in 2 3 4 6; do for c in 1 2; do for tile in 2048 4096; do
  if [ $cw = 16 ] && [ $c = 2 ]; then continue; fi
  run v1_w${cw}_s${st}_c${c}_t${tile} SPBLAS_B200_CONSUMER_WARPS=$cw SPBLAS_B200_STAGES=$st SPBLAS_B200_CTAS_PER_SM=$c SPBLAS_B200_TILE_ITEMS=$tile
done; done; done; done
