#!/bin/bash
# usage: scripts/gpu_ncu.sh TAG "bench args" [kernel regex] -- one `ncu --set full` capture of the workload's kernel (after
# warm-up), summarised ON the box (the reports of a --set full capture with sources exceed what gpurun copies back),
# + the launch list of the same command
TAG=$1; ARGS=$2; KRE=${3:-k_warp_tiled}
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$KRE -s 3 -c 1 -f -o /tmp/${TAG}_prof python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-other-workloads $ARGS > gpurun_out/${TAG}_p.log 2>&1
tail -2 gpurun_out/${TAG}_p.log
python scripts/ncu_summary.py /tmp/${TAG}_prof.ncu-rep ${TAG} x 0 gpurun_out > /dev/null
python scripts/ncu_lines.py /tmp/${TAG}_prof.ncu-rep 60 > gpurun_out/${TAG}_lines.txt 2>&1
ncu -i /tmp/${TAG}_prof.ncu-rep --page details --csv > gpurun_out/${TAG}_details.csv 2>/dev/null
SZ=$(stat -c %s /tmp/${TAG}_prof.ncu-rep 2>/dev/null || echo 0)
[ "$SZ" -lt 20000000 ] && cp /tmp/${TAG}_prof.ncu-rep gpurun_out/
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-other-workloads $ARGS > gpurun_out/${TAG}_l.log 2>&1
grep -c k_warp_tiled gpurun_out/${TAG}_launches.csv
head -20 gpurun_out/${TAG}_ncu_summary.txt
