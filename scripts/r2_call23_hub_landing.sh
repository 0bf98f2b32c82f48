#!/bin/bash
# The hub kernel with cp.async landing zones (spmv_hubl_stream_kernel, SPBLAS_B200_HUB_LANDING=1):
# parity (the hub suites with the env set), then timing against the LDG hub kernel on C4 (fp32)
# and R-MAT scale 24 fp64, over table sizes and warps per CTA (hw24/hw32 builds).
# Kept as the record of how profiles/r02_hub_landing_negative.jsonl was produced: the kernel, its
# environment switch and the hw24/hw32 builds were REMOVED after this run (slower at every shape;
# DESIGN 4.13) — the script no longer runs against the current library.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
out=gpurun_out/r2_hub_landing.jsonl; : > $out
echo "== parity with landing zones"
SPBLAS_B200_HUB_LANDING=1 timeout 600 python -m pytest tests/test_gpu_zhub.py tests/test_gpu_axpby.py -x -q -m gpu -k "not default_capacity" 2>&1 | tail -3
run() { EXP_VARIANT=3 "$@" >> $out 2>> gpurun_out/r2_hub_landing.err; }
run env EXP_HUB_COLS=32768 timeout 200 python scripts/exp_r2.py spmv c4 30
for h in 24576 32768 39936; do
  run env SPBLAS_B200_HUB_LANDING=1 EXP_HUB_COLS=$h timeout 200 python scripts/exp_r2.py spmv c4 30
done
for lib in hw24 hw32; do
  run env SPBLAS_B200_HUB_LANDING=1 SPBLAS_B200_LIB=$PWD/spblas_reference_b200/libspblas_b200_$lib.so EXP_HUB_COLS=49152 timeout 200 python scripts/exp_r2.py spmv c4 30
done
run env EXP_HUB_COLS=8192 timeout 200 python scripts/exp_r2.py spmv c5s24 30
for h in 8192 12288 19456; do
  run env SPBLAS_B200_HUB_LANDING=1 EXP_HUB_COLS=$h timeout 200 python scripts/exp_r2.py spmv c5s24 30
done
python - <<'PY'
import json
for l in open("gpurun_out/r2_hub_landing.jsonl"):
    if l.startswith("{"):
        d = json.loads(l); print(d["workload"], d["lib"], d["hub_count"], d["hub_ref_share"], "ms", d["ms"], "checksum", d["checksum"])
PY
tail -3 gpurun_out/r2_hub_landing.err
