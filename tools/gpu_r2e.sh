#!/bin/bash
mkdir -p gpurun_out
for d in 0 8 16 24 64 80 32 56 112; do echo "dbg $d"; PMMH_GRID_DEBUG=$d timeout 200 python tools/phase_clocks_grid.py 20 300 2>&1 | grep -v "zero\|shift" | tee -a gpurun_out/r2e_dbg.log; done
