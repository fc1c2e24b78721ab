#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sv_grid.py -x -q -m gpu > gpurun_out/r2g_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/r2g_tests.log
grep -E "passed|failed|exit|Error|assert" gpurun_out/r2g_tests.log | head -20
for ln in 20 19 18; do echo "logN $ln"; timeout 200 python tools/phase_clocks_grid.py $ln 300 2>&1 | tee -a gpurun_out/r2g_clocks.log; done
for d in 8 16 24; do echo "dbg $d"; PMMH_GRID_DEBUG=$d timeout 200 python tools/phase_clocks_grid.py 20 300 2>&1 | grep -v "wait\|zero" | tee -a gpurun_out/r2g_dbg.log; done
timeout 900 ncu --set full --import-source on --clock-control none -k regex:sv_grid_kernel -c 1 -f -o gpurun_out/r2g_grid python tools/probe_alg.py 6 20 100 1 > gpurun_out/r2g_ncu.log 2>&1
tail -2 gpurun_out/r2g_ncu.log
