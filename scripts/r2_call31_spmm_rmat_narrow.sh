#!/bin/bash
# SpMM on a power-law matrix with a NARROW B: group-per-row kernel (forced 0) vs merge-path
# ring kernel (forced 1) vs the default choice — should the row-length histogram steer it?
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
out=gpurun_out/r2_spmm_rmat_narrow.jsonl; : > $out
for cfg in "22 8 fp32" "22 16 fp32" "22 32 fp32" "22 4 fp64" "22 16 fp64"; do
  timeout 200 python scripts/exp_r2.py spmm_rmat $cfg >> $out 2>> gpurun_out/r2_spmm_rmat_narrow.err
done
python - <<'PY'
import json
for l in open("gpurun_out/r2_spmm_rmat_narrow.jsonl"):
    if l.startswith("{"):
        d = json.loads(l); print(d["dtype"], "k", d["k"], "forced", d["forced"], "variant", d["variant"], "ms", d["ms"], "maxrow", d["max_row_len"], "diff", d["max_abs_diff_vs_first"])
PY
tail -3 gpurun_out/r2_spmm_rmat_narrow.err
