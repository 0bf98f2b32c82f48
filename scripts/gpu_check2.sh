#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
for w in c1 c3k32; do timeout 600 python bench.py --workload $w --steps 20 --warmup 5 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; tail -2 gpurun_out/bench_$w.err; python -c "
import json
d=json.loads(open('gpurun_out/bench_$w.json').read().strip().splitlines()[-1]); print('$w', 'ms', round(d['ms_per_step'],4), 'cpu', d.get('cpu_baseline'))
"; done
