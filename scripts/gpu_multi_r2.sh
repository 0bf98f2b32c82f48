#!/bin/bash
# gpurun --gpus N -- 'bash scripts/gpu_multi_r2.sh N': the multi-GPU parity tests that fit N
# devices, then the bench as the driver launches it (C2 weak-scaled headline + C5 scale 27
# strong-scaled block with parity), fused-vs-NCCL decided by the set-up calibration.
N=${1:-2}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== multi-GPU parity tests (world size $N only: the smaller ones ran on the smaller boxes)"
SEL="$N"; [ "$N" = "2" ] && SEL="2 or timeout"
if [ "${SKIP_TESTS:-0}" != "1" ]; then
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -k "$SEL" > gpurun_out/r2_multi_tests_n$N.log 2>&1
tail -4 gpurun_out/r2_multi_tests_n$N.log
fi
echo "== where the exchange's cost goes (C2 weak)"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 scripts/exp_exchange.py 2>&1 | tail -1 | tee gpurun_out/r2_exchange_n$N.json
echo "== bench --gpus $N"
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
  --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 \
  > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
tail -c 1500 gpurun_out/r2_bench_n$N.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2_bench_n$N.json").read().strip().splitlines()[-1])
    print("C2 weak: value", round(d["value"], 1), "ms", round(d["ms_per_step"], 4), "kernel", round(d["roofline"]["kernel_ms"], 4),
          d["config"]["exchange_impl"][:40], d["config"]["exchange_calibration"], "parity", d["parity"]["pass"], d["parity"]["max_err_over_tol"])
    for k, b in d["configs"].items():
        print(k, "ms", round(b["ms"], 3), "gflops", round(b["gflops"], 1), "kernel_only", b.get("kernel_only_ms"), "variant", b.get("spmv_variant", b.get("spmm_variant")),
              "parity", b["parity"]["pass"], round(b["parity"]["max_err_over_tol"], 3), "exch", (b.get("exchange") or {}).get("impl", "")[:50], (b.get("exchange") or {}).get("calibration"), "wall", round(b["wall_s"], 1))
    print("all pass:", d["parity_all_pass"])
except Exception as e:
    print("parse failed", e)
PY
