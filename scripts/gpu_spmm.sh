#!/bin/bash
# SpMM: parity of both kernels, then C3 with the group kernel, the stream kernel and L2 fractions
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_spmm.py -m gpu -x -q > gpurun_out/pytest_spmm.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_spmm.log
tail -15 gpurun_out/pytest_spmm.log
run() { name=$1; wl=$2; shift; shift
  env "$@" timeout 300 python bench.py --workload $wl --steps 20 --warmup 5 2>gpurun_out/spmm_$name.err | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('$name', '$wl', 'ms=%.4f'%d['ms_per_step'], 'GFLOP/s=%.0f'%d['value'], 'frac=%.3f'%d['roofline']['frac'])
" | tee -a gpurun_out/spmm_sweep.txt
}
: > gpurun_out/spmm_sweep.txt
for wl in c3k32 c3k128; do
  run group $wl SPBLAS_B200_SPMM_VARIANT=0
  run flat_auto $wl SPBLAS_B200_SPMM_VARIANT=1
  for f in 0.0 0.2 0.35 0.5 1.0; do run flat_f$f $wl SPBLAS_B200_SPMM_VARIANT=1 SPBLAS_B200_SPMM_L2FRAC=$f; done
done
