#!/bin/bash
# usage: gpu_san.sh memcheck|racecheck|synccheck
mkdir -p gpurun_out
TOOL=$1
timeout 600 python tools/sanitize_small.py 6 2 3 5 1 > gpurun_out/r2_san_plain.log 2>&1 && \
timeout 1500 compute-sanitizer --tool $TOOL --print-limit 20 python tools/sanitize_small.py 6 2 3 5 1 > gpurun_out/r2_sanitizer_$TOOL.log 2>&1
echo "exit $?"; tail -12 gpurun_out/r2_sanitizer_$TOOL.log
