#!/bin/bash
# full evidence pass: parity, smoke, all benches (with cuSPARSE comparator + gather probe), reference arm,
# ncu launch list of the headline command, ncu --set full of the dominant kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench exit $?"
for w in c1 c4 c3k32 c3k128 c5 c1t c4t t1 t4; do timeout 600 python bench.py --workload $w --steps 30 --warmup 5 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; done
timeout 600 python bench.py --workload trsv --steps 10 --warmup 3 > gpurun_out/bench_trsv.json 2> gpurun_out/bench_trsv.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_c2.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
export SPBLAS_B200_NO_CUSPARSE=1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"spmv_pipe" -s 4 -c 1 -o gpurun_out/prof_c2 -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_c2.log 2>&1
for w in c4 c1; do
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"spmv_warp_stream" -s 2 -c 1 -o gpurun_out/prof_ws_$w -f python bench.py --workload $w --steps 3 --warmup 3 > gpurun_out/ncu_ws_$w.log 2>&1
done
cat gpurun_out/bench_c2.json
for w in c1 c4 c3k32 c3k128 c5 c1t c4t t1 t4 trsv ref; do python -c "
import json
d=json.loads(open('gpurun_out/bench_$w.json').read().strip().splitlines()[-1]); g=(d.get('roofline') or {}).get('gather_ceiling') or {}; print('$w', round(d['value'],1), d['unit'], 'ms', round(d['ms_per_step'],4), 'frac', d.get('roofline',{}).get('frac'), 'of_probe', g.get('frac_of_probe'), 'cusparse', (d.get('cusparse') or {}).get('ours_over_best_cusparse'))
"; done
