#!/bin/bash
# first GPU run of the grid kernel: parity tests, sanitizer on a small shape, timing + phase clocks
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sv_grid.py -x -q -m gpu > gpurun_out/r2a_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r2a_tests.log
tail -30 gpurun_out/r2a_tests.log
timeout 300 python tools/phase_clocks_grid.py 20 1000 > gpurun_out/r2a_clocks_20.log 2>&1; tail -20 gpurun_out/r2a_clocks_20.log
timeout 200 python tools/phase_clocks_grid.py 18 1000 > gpurun_out/r2a_clocks_18.log 2>&1; tail -3 gpurun_out/r2a_clocks_18.log
for alg in 6 5; do timeout 300 python tools/probe_alg.py $alg 20 1000 3 >> gpurun_out/r2a_probe.log 2>&1; done
tail -8 gpurun_out/r2a_probe.log
timeout 400 compute-sanitizer --tool memcheck python tools/sanitize_small.py 6 > gpurun_out/r2a_memcheck_grid.log 2>&1; tail -5 gpurun_out/r2a_memcheck_grid.log
