#!/bin/bash
# ncu --set full of the hub-stream kernel (32768-column table) and of the warp-stream kernel
# on the same R-MAT matrix (C4's: scale 24, fp32), one launch each after the warm-ups.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
S=${HUB_SCALE:-24}
timeout ${T1:-45} ncu --set full --clock-control none --import-source on -k regex:"spmv_hub_stream" -s 2 -c 1 -o gpurun_out/prof_hub_s$S -f python scripts/hub_ab.py $S fp32 32768 > gpurun_out/ncu_hub_s$S.log 2>&1
tail -2 gpurun_out/ncu_hub_s$S.log
timeout ${T2:-45} ncu --set full --clock-control none --import-source on -k regex:"spmv_warp_stream" -s 2 -c 1 -o gpurun_out/prof_ws_s$S -f python scripts/hub_ab.py $S fp32 0 > gpurun_out/ncu_ws_s$S.log 2>&1
tail -2 gpurun_out/ncu_ws_s$S.log
ls -la gpurun_out/*.ncu-rep
