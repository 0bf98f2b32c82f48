#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== tests of the changed code"
timeout 600 python -m pytest tests/test_gpu_spmm.py tests/test_gpu_zhub.py tests/test_gpu_trsv.py -x -q > gpurun_out/r2_call2_tests.log 2>&1
tail -3 gpurun_out/r2_call2_tests.log
echo "== SpMM column slicing"
timeout 300 python scripts/exp_r2.py spmm 32 -1,8,16,0 > gpurun_out/r2_spmm_slice_k32.jsonl 2>&1; cat gpurun_out/r2_spmm_slice_k32.jsonl
timeout 300 python scripts/exp_r2.py spmm 128 -1,8,16,32,0 > gpurun_out/r2_spmm_slice_k128.jsonl 2>&1; cat gpurun_out/r2_spmm_slice_k128.jsonl
echo "== L2 policy of the warp-stream walk"
for wl in c1 c4 c5s24 c5shard; do
  timeout 400 python scripts/exp_r2.py libs $wl base,ef,efel > gpurun_out/r2_l2_policy_$wl.jsonl 2>&1; cat gpurun_out/r2_l2_policy_$wl.jsonl
done
