#!/bin/bash
# usage: scripts/gpu_ncu_modes.sh TAG -- one ncu --set full capture of the warp kernel for each secondary workload
TAG=$1
mkdir -p gpurun_out
for WL in ${2:-5k7_lut_linear 8k_cubic_fixed 8k_lanczos4_fixed 8k_cubic_auto}; do
  timeout 300 ncu --set full --clock-control none -k regex:k_warp_tiled -s 3 -c 1 -o gpurun_out/${TAG}_${WL} python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --workload $WL > gpurun_out/${TAG}_${WL}.log 2>&1
  tail -1 gpurun_out/${TAG}_${WL}.log | cut -c1-120
done
