#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sv_grid.py -x -q -m gpu > gpurun_out/r2k_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/r2k_tests.log
grep -E "passed|failed|exit|Error|assert" gpurun_out/r2k_tests.log | head -20
for mb in default 0 40 96; do echo "persist $mb"; if [ $mb != default ]; then export PMMH_GRID_L2_PERSIST_MB=$mb; fi; timeout 200 python tools/phase_clocks_grid.py 20 300 2>&1 | grep -v "zero" | tee -a gpurun_out/r2k_clocks.log; done
unset PMMH_GRID_L2_PERSIST_MB
echo "dbg 32 (bulk P prefetch)"; PMMH_GRID_DEBUG=32 timeout 200 python tools/phase_clocks_grid.py 20 300 2>&1 | grep -v "zero\|wait" | tee -a gpurun_out/r2k_clocks.log
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum --clock-control none -k regex:sv_grid_kernel -c 1 python tools/probe_alg.py 6 20 100 1 2>&1 | grep -E "dram__|lts__|gpu__time" | tee gpurun_out/r2k_dram.log
