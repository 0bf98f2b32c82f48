#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_spmv.py tests/test_gpu_spmm.py tests/test_gpu_cpp_dropin.py -m gpu -x -q -k "matrix_opt or csc or dropin or reference" > gpurun_out/pytest_opt.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_opt.log
tail -8 gpurun_out/pytest_opt.log
for w in c1t c4t; do for mo in 1 0; do SPBLAS_B200_MATRIX_OPT=$mo timeout 300 python bench.py --workload $w --steps 20 --warmup 3 > gpurun_out/bench_${w}_opt$mo.json 2> gpurun_out/bench_${w}_opt$mo.err; tail -2 gpurun_out/bench_${w}_opt$mo.err; python -c "
import json
d=json.loads(open('gpurun_out/bench_${w}_opt$mo.json').read().strip().splitlines()[-1]); c=d.get('cusparse') or {}; print('$w matrix_opt=$mo', 'ms', round(d['ms_per_step'],4), 'frac', round(d['roofline']['frac'],3), 'inspect_ms', round(d['config']['inspect_ms'],2), 'vs cusparse^T', c.get('ours_over_best_cusparse'), c.get('CUSPARSE_SPMV_ALG_DEFAULT'))
"; done; done
