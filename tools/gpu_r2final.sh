#!/bin/bash
# final records of round 2: phase clocks, launch list of the bench command, ncu --set full of the grid kernel
mkdir -p gpurun_out
timeout 200 python tools/phase_clocks_grid.py 20 300 > gpurun_out/r2f_clocks.log 2>&1; head -3 gpurun_out/r2f_clocks.log | cut -c1-200
CMD="python bench.py --steps 2 --warmup 1 --no-configs --no-split --no-cpu-baseline --parity-steps 0 --e2e-steps 1"
$CMD > gpurun_out/r2f_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2f_launches.csv $CMD > gpurun_out/r2f_ncu1.log 2>&1
echo "launch list exit $?"
python tools/probe_alg.py 0 20 1000 1 > gpurun_out/r2f_plain2.log 2>&1 && \
ncu --set full --import-source on --clock-control none -k regex:sv_grid_kernel -c 1 -f -o gpurun_out/r2f_grid_T1000 python tools/probe_alg.py 0 20 1000 1 > gpurun_out/r2f_ncu2.log 2>&1
echo "full capture exit $?"; tail -1 gpurun_out/r2f_plain2.log | cut -c1-200
