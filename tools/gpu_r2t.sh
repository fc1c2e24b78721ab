#!/bin/bash
# launch list of the bench command + one full capture of the headline kernel (T=1000, N=2^20)
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --no-configs --no-split --no-cpu-baseline --parity-steps 0 --e2e-steps 1"
$CMD > gpurun_out/r2t_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2t_launches.csv $CMD > gpurun_out/r2t_ncu1.log 2>&1
echo "launch list exit $?"
python tools/probe_alg.py 0 20 1000 1 > gpurun_out/r2t_plain2.log 2>&1 && \
ncu --set full --import-source on --clock-control none -k regex:sv_grid_kernel -c 1 -f -o gpurun_out/r2t_grid_T1000 python tools/probe_alg.py 0 20 1000 1 > gpurun_out/r2t_ncu2.log 2>&1
echo "full capture exit $?"; tail -3 gpurun_out/r2t_ncu2.log; cat gpurun_out/r2t_plain2.log | tail -2
