#!/bin/bash
# e2e (host rvs) under different copy schedules of the host-streamed path
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-configs --no-split --no-cpu-baseline --parity-steps 0 --e2e-steps 4"
for v in "PMMH_GRID_U_TAPER=48" "PMMH_GRID_U_TAPER=32" "PMMH_GRID_U_CHUNK=512 PMMH_GRID_U_TAPER=48" "PMMH_GRID_U_CHUNK=1024 PMMH_GRID_U_TAPER=48"; do
  echo "== $v" | tee -a gpurun_out/r2_e2e_variants.log
  env PMMH_STREAM_TIMING=1 $v $CMD 2>&1 | grep "stream timing" | tail -1 | tee -a gpurun_out/r2_e2e_variants.log
  env $v $CMD 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('e2e', l['e2e']['value'], 'ms', 1048576*1000/l['e2e']['value']*1e3, 'e2e_ok', l['e2e_ok'])
" | tee -a gpurun_out/r2_e2e_variants.log
done
timeout 600 python -m pytest tests/test_gpu_sv_grid.py tests/test_gpu_estimators.py -x -q -k "streamed or host" 2>&1 | tail -2
