#!/bin/bash
# shared-memory hub table size with the branch-free load phase (8 gathers in flight per lane
# instead of 4: does the optimum move to smaller tables = larger L1?)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
out=gpurun_out/r2_hub_sweep_flat.jsonl; : > $out
for h in 8192 16384 24576 32768 40960; do
  EXP_VARIANT=3 EXP_HUB_COLS=$h timeout 200 python scripts/exp_r2.py spmv c4 30 >> $out 2>> gpurun_out/r2_hub_sweep_flat.err
done
for h in 4096 8192 12288 16384; do
  EXP_VARIANT=3 EXP_HUB_COLS=$h timeout 200 python scripts/exp_r2.py spmv c5s24 30 >> $out 2>> gpurun_out/r2_hub_sweep_flat.err
done
python - <<'PY'
import json
for l in open("gpurun_out/r2_hub_sweep_flat.jsonl"):
    if l.startswith("{"):
        d = json.loads(l); print(d["workload"], d["hub_count"], d["hub_ref_share"], "ms", d["ms"])
PY
tail -3 gpurun_out/r2_hub_sweep_flat.err
