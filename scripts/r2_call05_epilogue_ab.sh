#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== tests touched since the last call"
timeout 900 python -m pytest tests/test_gpu_axpby.py tests/test_gpu_trsv.py tests/test_gpu_spmv.py tests/test_gpu_zhub.py tests/test_gpu_cpp_dropin.py tests/test_gpu_host_exec.py -x -q > gpurun_out/r2_call5_pytest.log 2>&1
tail -12 gpurun_out/r2_call5_pytest.log
echo "== C2 headline: what slowed it down (x0 = ones vs U[0,1); kernels with / without the epilogue)"
for x0 in ones uniform; do for lib in base noepi; do
  L=""; [ $lib = noepi ] && L=$PWD/spblas_reference_b200/libspblas_b200_noepi.so
  SPBLAS_B200_BENCH_X0=$x0 SPBLAS_B200_LIB=$L timeout 300 python bench.py --configs none --no-cpu-baseline --no-e2e > gpurun_out/r2_c2_ab_${x0}_${lib}.json 2> gpurun_out/r2_c2_ab_${x0}_${lib}.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2_c2_ab_${x0}_${lib}.json").read().strip().splitlines()[-1])
    print("${x0} ${lib}: step", round(d["ms_per_step"], 4), "kernel", round(d["roofline"]["kernel_ms"], 4), "frac", round(d["roofline"]["frac"], 3), d["clocks"])
except Exception as e:
    print("${x0} ${lib}: failed", e); print(open("gpurun_out/r2_c2_ab_${x0}_${lib}.err").read()[-800:])
PY
done; done
echo "== C1 through the no-info overload (structure cache)"
timeout 300 python bench.py --configs c1 --no-e2e > gpurun_out/r2_bench_c1.json 2> gpurun_out/r2_bench_c1.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench_c1.json").read().strip().splitlines()[-1])
b = d["configs"]["c1"]; print("c1 ms", b["ms"], "no-info", b["no_info_overload"], b["parity"])
PY
echo "== c5shard through matrix_opt"
EXP_MATRIX_OPT=1 timeout 400 python scripts/exp_r2.py spmv c5shard 20
