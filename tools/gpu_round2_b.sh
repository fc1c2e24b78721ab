#!/bin/bash
mkdir -p gpurun_out
for d in 0 1 2 4 7; do echo "dbg $d" >> gpurun_out/r2b_dbg.log; PMMH_GRID_DEBUG=$d timeout 200 python tools/phase_clocks_grid.py 20 300 >> gpurun_out/r2b_dbg.log 2>&1; done
cat gpurun_out/r2b_dbg.log | grep -v "^  wait\|zero\|shift" 
timeout 900 ncu --set full --import-source on --clock-control none -k regex:sv_grid_kernel -c 1 -f -o gpurun_out/r2b_grid python tools/probe_alg.py 6 20 100 1 > gpurun_out/r2b_ncu.log 2>&1
tail -3 gpurun_out/r2b_ncu.log; ls -la gpurun_out/*.ncu-rep
