#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sv_exchange.py tests/test_gpu_sv_flps.py tests/test_gpu_model_hooks.py tests/test_gpu_host_exports.py -x -q -m gpu > gpurun_out/r2x_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/r2x_tests.log
tail -25 gpurun_out/r2x_tests.log
timeout 300 python tools/bench_aux.py chains 2>&1 | tail -3
