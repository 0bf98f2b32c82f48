#!/bin/bash
# ncu --set full of the warp-stream SpMV kernel on C4 and C1, and of the gather probe on C4
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export SPBLAS_B200_NO_CUSPARSE=1
for w in c4 c1; do
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"spmv_warp_stream" -s 2 -c 1 -o gpurun_out/prof_ws_$w -f python bench.py --workload $w --steps 3 --warmup 3 > gpurun_out/ncu_ws_$w.log 2>&1
tail -2 gpurun_out/ncu_ws_$w.log
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gather_probe" -s 4 -c 1 -o gpurun_out/prof_probe_c4 -f python bench.py --workload c4 --steps 3 --warmup 3 > gpurun_out/ncu_probe_c4.log 2>&1
ls -la gpurun_out/*.ncu-rep
