#!/bin/bash
# compile-time variants of the grid kernel (gpurun_variants/lib_<name>.so), timed back to back: probe (3 reps) + phase clocks
mkdir -p gpurun_out
cp pmmh-qn_b200/lib/libpmmh_qn_b200.so /tmp/lib_keep.so
for f in gpurun_variants/lib_*.so; do
  v=$(basename $f .so); v=${v#lib_}
  cp $f pmmh-qn_b200/lib/libpmmh_qn_b200.so
  echo "variant $v" | tee -a gpurun_out/r2_variants.log
  timeout 300 python tools/probe_alg.py 6 20 1000 3 2>&1 | tail -2 | cut -c1-130 | tee -a gpurun_out/r2_variants.log
  timeout 300 python tools/phase_clocks_grid.py 20 300 2>&1 | grep -E "^  A2" | awk '{printf "%s=%s ", $1, $(NF-2)}' | tee -a gpurun_out/r2_variants.log; echo | tee -a gpurun_out/r2_variants.log
done
cp /tmp/lib_keep.so pmmh-qn_b200/lib/libpmmh_qn_b200.so
