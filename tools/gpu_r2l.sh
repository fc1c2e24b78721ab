#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sv_grid.py -x -q -m gpu > gpurun_out/r2l_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/r2l_tests.log
grep -E "passed|failed|exit|Error|assert" gpurun_out/r2l_tests.log | head -20
timeout 200 python tools/phase_clocks_grid.py 20 300 2>&1 | tee -a gpurun_out/r2l_clocks.log
for d in 1 4 5; do echo "dbg $d"; PMMH_GRID_DEBUG=$d timeout 200 python tools/phase_clocks_grid.py 20 300 2>&1 | grep -v "wait\|zero" | tee -a gpurun_out/r2l_clocks.log; done
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum --clock-control none -k regex:sv_grid_kernel -c 1 python tools/probe_alg.py 6 20 100 1 2>&1 | grep -E "dram__|lts__|gpu__time" | tee gpurun_out/r2l_dram.log
