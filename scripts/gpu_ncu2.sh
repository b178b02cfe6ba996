#!/bin/bash
# usage: scripts/gpu_ncu2.sh TAG "bench args A" "bench args B" -- ncu --set full summaries of k_warp_tiled for two workloads (A/B comparison)
TAG=$1; shift
mkdir -p gpurun_out
i=0
for ARGS in "$@"; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_warp_tiled -s 3 -c 1 -f -o /tmp/${TAG}${i}_prof python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-other-workloads $ARGS > gpurun_out/${TAG}${i}_p.log 2>&1
  tail -1 gpurun_out/${TAG}${i}_p.log
  python scripts/ncu_summary.py /tmp/${TAG}${i}_prof.ncu-rep ${TAG}${i} x 0 gpurun_out > /dev/null
  python scripts/ncu_lines.py /tmp/${TAG}${i}_prof.ncu-rep 80 > gpurun_out/${TAG}${i}_lines.txt 2>&1
  i=$((i+1))
done
paste -d'|' gpurun_out/${TAG}0_ncu_summary.txt gpurun_out/${TAG}1_ncu_summary.txt | cut -c1-300
