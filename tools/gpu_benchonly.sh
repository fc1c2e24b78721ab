#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python bench.py > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench_1gpu.err; echo "bench exit $?"
python - <<'PY'
import json
l=json.loads(open('gpurun_out/r2_bench_1gpu.json').read().strip().splitlines()[-1])
for k in ("value","ms_per_step","e2e","e2e_device_rvs","gpu_launches","clocks"):
    print(k, json.dumps(l.get(k))[:300])
print("roofline", l["roofline"]["frac"], l["roofline"]["traffic"])
print("parity", {k: l["parity"][k] for k in ("generations_mismatched","ll_rel","grad_rel","near_ties","soft_ties")})
print("config4", l["config4_chains"]["gradient"]["value"], l["config4_chains"]["hessian"]["value"], l["config4_chains"]["hessian"]["kernel"])
PY
