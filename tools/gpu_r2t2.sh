#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sv_grid.py -x -q -m gpu > gpurun_out/r2t2_tests.log 2>&1; echo "tests exit $?"; tail -1 gpurun_out/r2t2_tests.log
timeout 200 python tools/phase_clocks_grid.py 20 300 > gpurun_out/r2t2_clocks.log 2>&1; cat gpurun_out/r2t2_clocks.log
CMD="python bench.py --steps 2 --warmup 1 --no-configs --no-split --no-cpu-baseline --parity-steps 0 --e2e-steps 1"
$CMD > gpurun_out/r2t2_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2t2_launches.csv $CMD > gpurun_out/r2t2_ncu1.log 2>&1
echo "launch list exit $?"
python tools/probe_alg.py 0 20 1000 1 > gpurun_out/r2t2_plain2.log 2>&1 && \
ncu --set full --import-source on --clock-control none -k regex:sv_grid_kernel -c 1 -f -o gpurun_out/r2t2_grid_T1000 python tools/probe_alg.py 0 20 1000 1 > gpurun_out/r2t2_ncu2.log 2>&1
echo "full capture exit $?"; tail -1 gpurun_out/r2t2_plain2.log
