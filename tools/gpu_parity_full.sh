#!/bin/bash
mkdir -p gpurun_out
free -g | head -2; nproc
timeout 2400 python tools/parity_full.py 20 1000 > gpurun_out/r2_parity_full_N2^20_T1000.json 2> gpurun_out/r2_parity_full.err; echo "exit $?"
cat gpurun_out/r2_parity_full_N2^20_T1000.json; tail -3 gpurun_out/r2_parity_full.err
