#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err; echo "bench exit $?"
tail -c 600 gpurun_out/r2_bench_${N}gpu.err
python - <<PY
import json
l=json.loads(open('gpurun_out/r2_bench_${N}gpu.json').read().strip().splitlines()[-1])
for k in ("value","ms_per_step","n_gpus","e2e","e2e_device_rvs","config3_subsampling","config4_chains","config5_split_pf"):
    print(k, json.dumps(l.get(k))[:700])
PY
