#!/bin/bash
# split particle filter with pack_direct: in-process ranks on one GPU (tests), then parity + throughput over NCCL
mkdir -p gpurun_out
N=${1:-2}
timeout 900 python -m pytest tests/test_gpu_sv_split.py -x -q 2>&1 | tail -3 | tee gpurun_out/r2split_tests.log
if [ "$N" -gt 1 ]; then
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tools/run_split_dist.py 24 --T 200 --out gpurun_out/r2_split_pf_${N}gpu_nccl_direct.jsonl 2>&1 | grep -v "^\*\*\*\|OMP_NUM" | tail -12 | cut -c1-400
fi
