#!/bin/bash
# y = A^T x through transposed(a) (inspect-built image) against cusparseSpMV(OPERATION_TRANSPOSE); transpose re-inspect timing
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for w in c1t c4t; do timeout 300 python bench.py --workload $w --steps 20 --warmup 3 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; tail -2 gpurun_out/bench_$w.err; python -c "
import json
d=json.loads(open('gpurun_out/bench_$w.json').read().strip().splitlines()[-1]); print('$w', 'ms', round(d['ms_per_step'],4), 'frac', round(d['roofline']['frac'],3), 'inspect_ms', round(d['config']['inspect_ms'],2), 'cusparse', d.get('cusparse'))
"; done
for w in t1 t4; do timeout 300 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; tail -2 gpurun_out/bench_$w.err; python -c "
import json
d=json.loads(open('gpurun_out/bench_$w.json').read().strip().splitlines()[-1]); print('$w', 'ms', round(d['ms_per_step'],4), 'GB/s', round(d['value'],1), 'inspect_ms', round(d['config']['inspect_ms'],3))
"; done
