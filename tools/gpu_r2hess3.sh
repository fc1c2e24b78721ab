#!/bin/bash
# Hessian instantiation of the grid kernel: phase clocks, then one ncu --set full capture (T = 100)
mkdir -p gpurun_out
PMMH_PROBE_HESS=1 timeout 300 python tools/phase_clocks_grid.py 20 300 > gpurun_out/r2hess_phase_clocks.txt 2>&1
PMMH_PROBE_HESS=1 python tools/probe_alg.py 6 20 100 2 > gpurun_out/r2hess3_plain.log 2>&1 && \
PMMH_PROBE_HESS=1 ncu --set full --import-source on --clock-control none -k regex:sv_grid_kernel -c 1 -f -o gpurun_out/r2hess3_grid_hess_T100 python tools/probe_alg.py 6 20 100 1 > gpurun_out/r2hess3_ncu.log 2>&1
echo "full capture exit $?"; tail -1 gpurun_out/r2hess3_plain.log
