#!/bin/bash
# Hub variant: the GPU tests with the L1-bypassing gathers on, then the A/B over table sizes on
# R-MAT scale 24 (C4's matrix).  Everything lands in gpurun_out/.
mkdir -p gpurun_out
SPBLAS_B200_HUB_GATHER_CG=1 SPBLAS_B200_RUN_UNVALIDATED=1 timeout ${HUB_TEST_TIMEOUT:-25} python -m pytest tests/test_gpu_zhub.py -x -q > gpurun_out/hub_tests_cg.log 2>&1
echo "pytest exit $?" >> gpurun_out/hub_tests_cg.log
tail -4 gpurun_out/hub_tests_cg.log
timeout ${HUB_AB24_TIMEOUT:-30} python scripts/hub_ab.py 24 fp32 0,0g,32768,32768g,40960g,49152g > gpurun_out/hub_ab_s24_fp32_cg.log 2>&1
echo "ab exit $?" >> gpurun_out/hub_ab_s24_fp32_cg.log
cat gpurun_out/hub_ab_s24_fp32_cg.log
