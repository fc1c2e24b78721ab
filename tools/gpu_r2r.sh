#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 3 --warmup 3 --split-steps 100 > gpurun_out/r2r_bench.json 2> gpurun_out/r2r_bench.err; echo "bench exit $?"
tail -c 600 gpurun_out/r2r_bench.err
python - <<'PY'
import json
l=json.loads(open('gpurun_out/r2r_bench.json').read().strip().splitlines()[-1])
for k,v in l.items():
    print(k, json.dumps(v)[:700])
PY
