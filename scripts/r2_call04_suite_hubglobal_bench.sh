#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== the whole GPU suite"
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_call4_pytest.log 2>&1
tail -15 gpurun_out/r2_call4_pytest.log
echo "== hub table in global memory"
(EXP_MATRIX_OPT=1 timeout 400 python scripts/exp_r2.py spmv c5shard 20
 timeout 400 python scripts/exp_r2.py spmv c5shard 20
 EXP_VARIANT=4 EXP_HUB_COLS=2000000 timeout 400 python scripts/exp_r2.py spmv c5shard 20
 EXP_VARIANT=4 EXP_HUB_COLS=16000000 timeout 400 python scripts/exp_r2.py spmv c5shard 20
 EXP_VARIANT=2 timeout 300 python scripts/exp_r2.py spmv c4 30
 EXP_VARIANT=3 timeout 300 python scripts/exp_r2.py spmv c4 30
 EXP_VARIANT=4 EXP_HUB_COLS=1000000 timeout 300 python scripts/exp_r2.py spmv c4 30
 EXP_VARIANT=4 EXP_HUB_COLS=4000000 timeout 300 python scripts/exp_r2.py spmv c4 30
 EXP_VARIANT=4 EXP_HUB_COLS=262144 timeout 300 python scripts/exp_r2.py spmv c4 30
 EXP_VARIANT=2 timeout 300 python scripts/exp_r2.py spmv c5s24 30
 EXP_VARIANT=3 timeout 300 python scripts/exp_r2.py spmv c5s24 30
 EXP_VARIANT=4 EXP_HUB_COLS=2000000 timeout 300 python scripts/exp_r2.py spmv c5s24 30
 EXP_VARIANT=2 timeout 300 python scripts/exp_r2.py spmv c1 30
) > gpurun_out/r2_hub_global.jsonl 2>&1
cat gpurun_out/r2_hub_global.jsonl | cut -c1-330
echo "== bench: headline + c1, c3k32, c4, and c5 at scale 24 (the code path)"
timeout 900 python bench.py --configs c1,c3k32,c4,c5 --c5-scale 24 > gpurun_out/r2_bench_debug.json 2> gpurun_out/r2_bench_debug.err
tail -c 600 gpurun_out/r2_bench_debug.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2_bench_debug.json").read().strip().splitlines()[-1])
    print("value", d["value"], "frac", d["roofline"]["frac"], "parity", d["parity"], "e2e", d["e2e"]["value"])
    for k, b in d["configs"].items():
        print(k, "ms", b["ms"], "frac", b["roofline"]["frac"], "parity", b["parity"]["max_err_over_tol"], b["parity"]["pass"], "cpu", (b.get("cpu_baseline") or {}).get("value"), "e2e", b["e2e"]["value"], {kk: vv for kk, vv in b.items() if kk in ("no_info_overload", "plain_operand", "spmv_variant", "kernel_only_ms")})
    print("errors", d.get("config_errors"))
except Exception as e:
    print("bench parse failed", e)
PY
