#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_host_exports.py tests/test_gpu_sv_grid.py tests/test_gpu_sv_exchange.py -x -q -m gpu > gpurun_out/r2s_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/r2s_tests.log
tail -15 gpurun_out/r2s_tests.log
timeout 1200 python bench.py --steps 3 --warmup 3 --split-steps 100 > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err; echo "bench exit $?"
tail -c 800 gpurun_out/r2s_bench.err
python - <<'PY'
import json
l=json.loads(open('gpurun_out/r2s_bench.json').read().strip().splitlines()[-1])
for k,v in l.items():
    print(k, json.dumps(v)[:900])
PY
