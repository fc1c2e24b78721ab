#!/bin/bash
mkdir -p gpurun_out
for d in 0 1 4 5 7; do
echo "dbg $d" >> gpurun_out/r2i_dram.log
PMMH_GRID_DEBUG=$d timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__t_bytes.sum,gpu__time_duration.sum --clock-control none -k regex:sv_grid_kernel -c 1 python tools/probe_alg.py 6 20 100 1 2>&1 | grep -E "dram__|lts__|gpu__time" >> gpurun_out/r2i_dram.log
done
cat gpurun_out/r2i_dram.log
