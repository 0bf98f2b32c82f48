#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== default bench, as the driver runs it (N=1, every config, C5 at scale 27)"
SECONDS=0; SPBLAS_B200_BENCH_VERBOSE=1 timeout 1500 python -X faulthandler bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
echo "wall seconds: $SECONDS"
tail -c 1500 gpurun_out/r2_bench_n1.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2_bench_n1.json").read().strip().splitlines()[-1])
    print("C2: value", round(d["value"], 1), "ms", round(d["ms_per_step"], 4), "kernel", round(d["roofline"]["kernel_ms"], 4), "frac", round(d["roofline"]["frac"], 3), "parity", d["parity"]["pass"], round(d["parity"]["max_err_over_tol"], 3), "e2e", round(d["e2e"]["value"], 1), "cpu", round(d["cpu_baseline"]["value"], 3))
    for k, b in d["configs"].items():
        print(k, "ms", round(b["ms"], 4), "gflops", round(b["gflops"], 1), "frac", round(b["roofline"]["frac"], 3), "kernel_only", b.get("kernel_only_ms"), "variant", b.get("spmv_variant", b.get("spmm_variant")),
              "parity", b["parity"]["pass"], round(b["parity"]["max_err_over_tol"], 3), "cpu", round((b.get("cpu_baseline") or {}).get("value", 0), 3), "e2e", round(b["e2e"]["value"], 1), "wall", round(b["wall_s"], 1),
              {kk: vv for kk, vv in b.items() if kk in ("no_info_overload", "plain_operand", "hub_columns", "hub_reference_share", "generate_s", "first_execute_ms")})
    print("all pass:", d["parity_all_pass"], "errors:", d.get("config_errors"))
except Exception as e:
    print("parse failed", e)
PY
nvidia-smi --query-gpu=memory.used --format=csv
