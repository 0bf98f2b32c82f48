#!/bin/bash
# warp-stream SpMV kernel (variant 2): parity, then C1 / C4 / C5(N=1) against the defaults, stream-length sweep
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_spmv.py -m gpu -x -q -k "every_kernel_variant" > gpurun_out/pytest_ws.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_ws.log
tail -15 gpurun_out/pytest_ws.log
run() { w=$1; name=$2; shift; shift
  env SPBLAS_B200_NO_CUSPARSE=1 "$@" timeout 300 python bench.py --workload $w --steps 30 --warmup 5 --no-cpu-baseline --no-e2e 2>gpurun_out/ws_${w}_$name.err | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('$w', '$name', 'ms=%.4f'%d['ms_per_step'], 'frac=%.3f'%d['roofline']['frac'], d['config'].get('kernel_only_ms',''))
" | tee -a gpurun_out/ws_sweep.txt
  tail -2 gpurun_out/ws_${w}_$name.err
}
: > gpurun_out/ws_sweep.txt
for w in c1 c4 c5; do
  run $w default
  run $w v2 SPBLAS_B200_SPMV_VARIANT=2
  for it in 1024 2048 8192 16384; do
    run $w v2_items$it SPBLAS_B200_SPMV_VARIANT=2 SPBLAS_B200_WS_ITEMS=$it
  done
done
run c2 v2 SPBLAS_B200_SPMV_VARIANT=2
for w in c3k32 c3k128; do timeout 600 python bench.py --workload $w --steps 20 --warmup 5 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; tail -2 gpurun_out/bench_$w.err; python -c "
import json
d=json.loads(open('gpurun_out/bench_$w.json').read().strip().splitlines()[-1]); print('$w', round(d['value'],1), d['unit'], 'ms', round(d['ms_per_step'],4), 'frac', d.get('roofline',{}).get('frac'), 'cusparse', d.get('cusparse'))
"; done
