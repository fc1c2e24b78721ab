#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_lockstep.py tests/test_gpu_sv_exchange.py tests/test_gpu_sv_split.py -x -q -m gpu 2>&1 | tail -2
timeout 1200 python bench.py --steps 2 --warmup 3 --no-split --no-cpu-baseline --parity-steps 0 --e2e-steps 1 > gpurun_out/r2lb_bench.json 2> gpurun_out/r2lb_bench.err; echo "bench exit $?"
python - <<'PY'
import json
l=json.loads(open('gpurun_out/r2lb_bench.json').read().strip().splitlines()[-1])
print(json.dumps(l["config4_chains"].get("lockstep_cpmh"))[:700])
PY
