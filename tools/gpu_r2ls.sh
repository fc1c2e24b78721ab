#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_lockstep.py tests/test_gpu_estimators.py -x -q -m gpu > gpurun_out/r2ls_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/r2ls_tests.log
tail -30 gpurun_out/r2ls_tests.log
