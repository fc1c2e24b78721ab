#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sv_grid.py -x -q -k "hessian" 2>&1 | tail -5 | tee gpurun_out/r2hess2_tests.log
for th in 1024 512; do echo "threads $th"; PMMH_GRID_THREADS=$th PMMH_PROBE_HESS=1 timeout 300 python tools/phase_clocks_grid.py 20 300 2>&1 | tee -a gpurun_out/r2hess2_clocks.log; done
