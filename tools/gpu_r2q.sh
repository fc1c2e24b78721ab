#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sv_grid.py -x -q -m gpu > gpurun_out/r2q_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/r2q_tests.log
grep -E "passed|failed|exit|Error|assert" gpurun_out/r2q_tests.log | head -20
timeout 200 python tools/phase_clocks_grid.py 20 300 2>&1 | tee gpurun_out/r2q_clocks.log
for d in 8 1 4 5; do echo "dbg $d"; PMMH_GRID_DEBUG=$d timeout 200 python tools/phase_clocks_grid.py 20 300 2>&1 | tee -a gpurun_out/r2q_clocks.log; done
