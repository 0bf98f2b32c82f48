#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_trsv.py -m gpu -x -q > gpurun_out/pytest_trsv.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_trsv.log
tail -6 gpurun_out/pytest_trsv.log
timeout 120 python bench.py --workload trsv --grid 1024 --steps 10 --warmup 3 > gpurun_out/bench_trsv_1024.json 2> gpurun_out/bench_trsv_1024.err; tail -2 gpurun_out/bench_trsv_1024.err
timeout 300 python bench.py --workload trsv --steps 10 --warmup 3 > gpurun_out/bench_trsv.json 2> gpurun_out/bench_trsv.err; tail -2 gpurun_out/bench_trsv.err
for f in trsv_1024 trsv; do python -c "
import json
d=json.loads(open('gpurun_out/bench_$f.json').read().strip().splitlines()[-1]); c=d['config']; print('$f', 'ms', round(d['ms_per_step'],3), 'direct', round(c['ms_per_step_level_by_level_launches'],3), 'levels', c['levels'], 'inspect_ms', round(c['inspect_ms'],1), 'sweeps', c['inspect_sweeps'], 'same', c['graph_and_direct_bit_identical'], 'cpu', d.get('cpu_baseline'))
"; done
