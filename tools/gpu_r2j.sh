#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:sv_grid_kernel -c 1 -f -o gpurun_out/r2j_grid python tools/probe_alg.py 6 20 100 1 > gpurun_out/r2j_ncu.log 2>&1
tail -2 gpurun_out/r2j_ncu.log
