#!/bin/bash
# parity (incl. the host-buffer execute), smoke, headline bench with pipelined e2e + cuSPARSE comparator,
# side workloads with the comparator
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench exit $?"
cat gpurun_out/bench_c2.json; tail -3 gpurun_out/bench_c2.err
for ch in 4 8 32; do SPBLAS_B200_HOST_CHUNKS=$ch SPBLAS_B200_NO_CUSPARSE=1 timeout 300 python bench.py --no-cpu-baseline --steps 10 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('chunks $ch e2e ms', d['e2e']['ms_per_step'], 'serial', d['e2e']['serial_ms_per_step'])
"; done
for w in c1 c4 c3k32 c3k128; do timeout 600 python bench.py --workload $w --steps 30 --warmup 5 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; tail -2 gpurun_out/bench_$w.err; python -c "
import json
d=json.loads(open('gpurun_out/bench_$w.json').read().strip().splitlines()[-1]); print('$w', round(d['value'],1), d['unit'], 'ms', round(d['ms_per_step'],4), 'frac', d.get('roofline',{}).get('frac'), 'cusparse', d.get('cusparse'))
"; done
