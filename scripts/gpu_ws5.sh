#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for co in -1 20 25 32 50; do
for w in c1 c4 c5; do SPBLAS_B200_WS_CARVEOUT=$co SPBLAS_B200_NO_CUSPARSE=1 timeout 600 python bench.py --workload $w --steps 30 --warmup 5 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; tail -2 gpurun_out/bench_$w.err; python -c "
import json
d=json.loads(open('gpurun_out/bench_$w.json').read().strip().splitlines()[-1]); g=d['roofline'].get('gather_ceiling') or {}; print('carveout $co', '$w', 'ms', round(d['ms_per_step'],4), 'frac', round(d['roofline']['frac'],3), 'probe_ms', g.get('probe_ms'), 'frac_of_probe', g.get('frac_of_probe'))
"; done; done
