#!/bin/bash
# A/B of the walk's branch-free load phase (B200_WS_FLAT) and its register budget
# (B200_WS_CTAS / _WIDE): old = round-2 production (FLAT=0, 6/5 CTAs per SM), base = the shipped build =
# FLAT=1 5/4, f65 = FLAT=1 6/5 (spills), f43 = FLAT=1 4/3.  The A/B builds are made with
#   make -C spblas_reference_b200/csrc OUT=../libspblas_b200_<name>.so OBJDIR=../../build/obj_<name> \
#        EXTRA="-DB200_WS_FLAT=<0|1> -DB200_WS_CTAS=<n> -DB200_WS_CTAS_WIDE=<n>"
# and are not kept in the tree (30 MB each).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
out=gpurun_out/r2_flat_walk_ab.jsonl; : > $out
echo "== parity of the new default build (SpMV suites)"
timeout 600 python -m pytest tests/test_gpu_spmv.py tests/test_gpu_zhub.py tests/test_gpu_axpby.py -x -q -m gpu 2>&1 | tail -3
for wl in c1 c4 c5s24; do
  timeout 300 python scripts/exp_r2.py libs $wl old,f65,base,f43 30 >> $out 2>> gpurun_out/r2_flat_walk_ab.err
done
for wl in c4 c5shard; do
  EXP_MATRIX_OPT=1 timeout 400 python scripts/exp_r2.py libs $wl old,f65,base,f43 30 >> $out 2>> gpurun_out/r2_flat_walk_ab.err
done
python - <<'PY'
import json
for l in open("gpurun_out/r2_flat_walk_ab.jsonl"):
    if l.startswith("{"):
        d = json.loads(l)
        print(d["workload"], d["lib"], "opt" if d["matrix_opt"] else "plain", "variant", d["variant"], "ms", d["ms"])
PY
tail -5 gpurun_out/r2_flat_walk_ab.err
