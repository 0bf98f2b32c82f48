#!/bin/bash
# parity + smoke + headline bench, then a shape sweep of the general-path SpMV kernels on C1 and C4
# (tile items x kernel variant x CTAs/SM) through the tuning environment variables (csrc/cabi.cu).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench exit $?"
cat gpurun_out/bench_c2.json
run() { w=$1; name=$2; shift; shift
  env "$@" timeout 300 python bench.py --workload $w --steps 30 --warmup 5 --no-cpu-baseline --no-e2e 2>gpurun_out/sweepg_${w}_$name.err | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('$w', '$name', 'ms=%.4f'%d['ms_per_step'], 'frac=%.3f'%d['roofline']['frac'])
" | tee -a gpurun_out/sweep_general.txt
}
: > gpurun_out/sweep_general.txt
for w in c1 c4; do
  run $w default
  for t in 1024 1536 3072 4096; do
    run $w v0_t$t SPBLAS_B200_SPMV_VARIANT=0 SPBLAS_B200_TILE_ITEMS=$t
  done
  for cfg in "8 3 2048 3" "8 4 1536 3" "8 5 1024 3" "16 2 2048 3" "16 2 1536 4" "16 1 4096 4"; do
    set -- $cfg
    run $w v1_w$1_c$2_t$3_s$4 SPBLAS_B200_SPMV_VARIANT=1 SPBLAS_B200_CONSUMER_WARPS=$1 SPBLAS_B200_CTAS_PER_SM=$2 SPBLAS_B200_TILE_ITEMS=$3 SPBLAS_B200_STAGES=$4
  done
done
