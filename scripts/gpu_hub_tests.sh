#!/bin/bash
# the hub variant's GPU tests alone
mkdir -p gpurun_out
SPBLAS_B200_RUN_UNVALIDATED=1 timeout ${HUB_TEST_TIMEOUT:-25} python -m pytest tests/test_gpu_zhub.py -x -q > gpurun_out/hub_tests_final.log 2>&1
echo "pytest exit $?" >> gpurun_out/hub_tests_final.log
tail -4 gpurun_out/hub_tests_final.log
