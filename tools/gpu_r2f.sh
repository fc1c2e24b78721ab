#!/bin/bash
mkdir -p gpurun_out
for th in 1024 512; do
PMMH_GRID_THREADS=$th timeout 600 python -m pytest tests/test_gpu_sv_grid.py -x -q -m gpu > gpurun_out/r2f_tests_$th.log 2>&1; echo "tests($th) exit $?" >> gpurun_out/r2f_tests_$th.log
grep -E "passed|failed|exit|Error|assert" gpurun_out/r2f_tests_$th.log | head -20
done
for th in 1024 512; do echo "threads $th"; PMMH_GRID_THREADS=$th timeout 200 python tools/phase_clocks_grid.py 20 300 2>&1 | tee -a gpurun_out/r2f_clocks.log; done
