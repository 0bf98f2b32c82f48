#!/bin/bash
# validation pass: parity, smoke, headline bench, ncu of the general-path SpMV kernels (C1, C4)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench exit $?"
cat gpurun_out/bench_c2.json
for w in c1 c4; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"spmv_merge_tile|spmv_pipe" -s 2 -c 1 -o gpurun_out/prof_$w -f python bench.py --workload $w --steps 3 --warmup 3 > gpurun_out/ncu_$w.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
