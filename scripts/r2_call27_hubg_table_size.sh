#!/bin/bash
# global hub table size on the scale-27 shard after the L1-priority tuning and the branch-free
# walk: ncu shows half of the table's references missing L2 with the 63 MB default (the two L2
# partitions each hold a copy of lines used by all SMs) — does a table that fits twice do better?
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
out=gpurun_out/r2_hubg_table_size.jsonl; : > $out
for h in 0 1000000 2000000 3000000 4000000 6000000; do
  SPBLAS_B200_HUB_COLS=$h EXP_MATRIX_OPT=1 timeout 300 python scripts/exp_r2.py spmv c5shard 20 >> $out 2>> gpurun_out/r2_hubg_table_size.err
done
python - <<'PY'
import json
for l in open("gpurun_out/r2_hubg_table_size.jsonl"):
    if l.startswith("{"):
        d = json.loads(l); print(d["workload"], "variant", d["variant"], d["hub_count"], d["hub_ref_share"], "ms", d["ms"])
PY
tail -3 gpurun_out/r2_hubg_table_size.err
