#!/bin/bash
# gather-ceiling probe next to the general-path SpMV workloads
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for w in c1 c4 c5; do SPBLAS_B200_NO_CUSPARSE=1 timeout 600 python bench.py --workload $w --steps 30 --warmup 5 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; tail -2 gpurun_out/bench_$w.err; python -c "
import json
d=json.loads(open('gpurun_out/bench_$w.json').read().strip().splitlines()[-1]); print('$w', 'ms', round(d['ms_per_step'],4), 'frac', round(d['roofline']['frac'],3), 'probe', d['roofline'].get('gather_ceiling'))
"; done
