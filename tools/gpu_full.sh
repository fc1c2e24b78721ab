#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2_gputests.log 2>&1; echo "tests exit $?"
tail -8 gpurun_out/r2_gputests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 1200 python bench.py > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench_1gpu.err; echo "bench exit $?"
tail -c 400 gpurun_out/r2_bench_1gpu.err
python - <<'PY'
import json
l=json.loads(open('gpurun_out/r2_bench_1gpu.json').read().strip().splitlines()[-1])
for k in ("value","ms_per_step","roofline","e2e","e2e_device_rvs","parity","gpu_launches","clocks","config5_split_pf","cpu_baseline_same_config"):
    print(k, json.dumps(l.get(k))[:600])
PY
