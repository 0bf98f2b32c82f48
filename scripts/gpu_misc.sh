#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for w in c1 c4; do timeout 600 python bench.py --workload $w --steps 30 --warmup 5 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; python -c "
import json
d=json.loads(open('gpurun_out/bench_$w.json').read().strip().splitlines()[-1]); print('$w', round(d['value'],1), d['unit'], 'ms', round(d['ms_per_step'],4), 'frac', d.get('roofline',{}).get('frac'))
"; done
timeout 900 python bench.py --workload c5 --steps 20 --warmup 3 > gpurun_out/bench_c5_n1.json 2> gpurun_out/bench_c5_n1.err; tail -c 1500 gpurun_out/bench_c5_n1.json; tail -3 gpurun_out/bench_c5_n1.err
# sanitizer passes over the small-shape parity tests (pipelined kernel + fallback)
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_spmv.py -x -q -k "vs_oracle or probe or inspect_structures or zero_sized or shard" > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck exit $?"; tail -4 gpurun_out/sanitizer_memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_spmv.py -x -q -k "probe or inspect_structures" > gpurun_out/sanitizer_racecheck.log 2>&1; echo "racecheck exit $?"; tail -4 gpurun_out/sanitizer_racecheck.log
