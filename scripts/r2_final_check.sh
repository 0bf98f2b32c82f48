#!/bin/bash
# What the driver runs at round end, on one GPU: smoke(), the GPU suite, the reference arm, the bench.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== pytest -m gpu"; SECONDS=0
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/r2_final_pytest.log 2>&1; tail -4 gpurun_out/r2_final_pytest.log; echo "pytest wall: $SECONDS s"
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 5 --warmup 2 | cut -c1-400
echo "== bench"; SECONDS=0
timeout 1500 python bench.py > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err; echo "bench wall: $SECONDS s"; tail -c 600 gpurun_out/r2_final_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_final_bench.json").read().strip().splitlines()[-1])
print("C2", round(d["value"], 1), round(d["roofline"]["frac"], 3), d["parity"]["pass"], "e2e", round(d["e2e"]["value"], 1))
for k, b in d["configs"].items():
    print(k, round(b["ms"], 4), round(b["roofline"]["frac"], 3), b["parity"]["pass"], round(b["parity"]["max_err_over_tol"], 3), b["roofline"]["traffic"], b.get("no_info_overload", {}).get("overhead_us"))
print("all pass", d["parity_all_pass"], d.get("config_errors"))
PY
