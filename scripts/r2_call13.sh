#!/bin/bash
# round 2, call 13: options/validation tests, legacy-MMA rate micro-benchmark
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 120 scripts/ubench/mma_tf32 > gpurun_out/r2o_mma_tf32.txt 2>&1; cat gpurun_out/r2o_mma_tf32.txt
timeout 900 python -m pytest tests/test_gpu_options.py tests/test_gpu_ccd.py tests/test_gpu_integer_tier.py tests/test_gpu_als.py -m gpu -q -x > gpurun_out/r2o_pytest.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/r2o_pytest.log
