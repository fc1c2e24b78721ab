#!/bin/bash
mkdir -p gpurun_out
for d in 0 4096 8192 12288; do echo "dbg $d"; PMMH_GRID_DEBUG=$d timeout 200 python tools/phase_clocks_grid.py 20 300 2>&1 | tee -a gpurun_out/r2z3_clocks.log | grep -E "us_per_step|A1:child|records|score|A2|rank" | cut -c1-100; done
