#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_spmv.py tests/test_gpu_host_exec.py -m gpu -x -q > gpurun_out/pytest_ws.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_ws.log
tail -4 gpurun_out/pytest_ws.log
for w in c1 c4 c5; do SPBLAS_B200_NO_CUSPARSE=1 timeout 600 python bench.py --workload $w --steps 30 --warmup 5 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; tail -2 gpurun_out/bench_$w.err; python -c "
import json
d=json.loads(open('gpurun_out/bench_$w.json').read().strip().splitlines()[-1]); g=d['roofline'].get('gather_ceiling') or {}; print('$w', 'ms', round(d['ms_per_step'],4), 'frac', round(d['roofline']['frac'],3), 'probe_ms', g.get('probe_ms'), 'frac_of_probe', g.get('frac_of_probe'))
"; done
