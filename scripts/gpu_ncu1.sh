#!/bin/bash
# usage: scripts/gpu_ncu1.sh TAG "bench args" [lines] -- ncu --set full summary + per-source-line split of k_warp_tiled for one workload
TAG=$1; ARGS=$2; N=${3:-150}
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_warp_tiled -s 3 -c 1 -f -o /tmp/${TAG}_prof python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-other-workloads $ARGS > gpurun_out/${TAG}_p.log 2>&1
tail -1 gpurun_out/${TAG}_p.log
python scripts/ncu_summary.py /tmp/${TAG}_prof.ncu-rep ${TAG} x 0 gpurun_out > /dev/null
python scripts/ncu_lines.py /tmp/${TAG}_prof.ncu-rep $N > gpurun_out/${TAG}_lines.txt 2>&1
head -16 gpurun_out/${TAG}_ncu_summary.txt
