#!/bin/bash
# memcheck / racecheck over the SpMV variants, the host-buffer execute and the transpose, then the
# transpose workloads, parity + smoke + headline bench exactly as the driver runs them
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
SEL="every_kernel_variant and mixed or host_execute_bit_identical and hub and 16 or transpose_vs_oracle and hub or reference_transpose_test"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_spmv.py tests/test_gpu_host_exec.py tests/test_gpu_transpose.py -m gpu -x -q -k "$SEL" > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck exit $?" >> gpurun_out/sanitizer_memcheck.log
tail -4 gpurun_out/sanitizer_memcheck.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_spmv.py -m gpu -x -q -k "every_kernel_variant and 2 and (mixed or hub) and 256" > gpurun_out/sanitizer_racecheck.log 2>&1; echo "racecheck exit $?" >> gpurun_out/sanitizer_racecheck.log
tail -4 gpurun_out/sanitizer_racecheck.log
for w in t1 t4; do timeout 300 python bench.py --workload $w --steps 20 --warmup 3 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; tail -2 gpurun_out/bench_$w.err; python -c "
import json
d=json.loads(open('gpurun_out/bench_$w.json').read().strip().splitlines()[-1]); print('$w', 'ms', round(d['ms_per_step'],4), 'GB/s', round(d['value'],1), 'inspect_ms', round(d['config']['inspect_ms'],3), 'cpu', d.get('cpu_baseline'))
"; done
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench exit $?"
cat gpurun_out/bench_c2.json
