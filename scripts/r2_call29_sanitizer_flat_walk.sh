#!/bin/bash
# compute-sanitizer memcheck over the walk kernels after the branch-free load phase (plain walk,
# shared-memory and global hub tables, the 4-argument product), racecheck over the plain walk
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
SEL="every_kernel_variant and mixed or lanes_outside or hub_structures_and_product and short or global_hub_table or axpby and spmv"
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_spmv.py tests/test_gpu_zhub.py tests/test_gpu_axpby.py -m gpu -x -q -k "$SEL" > gpurun_out/r2_sanitizer_memcheck_flat.log 2>&1; echo "memcheck exit $?" >> gpurun_out/r2_sanitizer_memcheck_flat.log
tail -4 gpurun_out/r2_sanitizer_memcheck_flat.log
timeout 400 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_spmv.py -m gpu -x -q -k "lanes_outside and 256 or every_kernel_variant and 2 and mixed and 256" > gpurun_out/r2_sanitizer_racecheck_flat.log 2>&1; echo "racecheck exit $?" >> gpurun_out/r2_sanitizer_racecheck_flat.log
tail -4 gpurun_out/r2_sanitizer_racecheck_flat.log
