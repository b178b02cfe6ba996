#!/bin/bash
# usage: scripts/gpu_final_ncu.sh TAG -- folded-chain tests, then ncu --set full of the headline kernel and of the analytic 4K pair, and the launch list of the default bench command
TAG=$1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fast_chain.py -x -q 2>&1 | tail -3
for W in 8k_rot_poly_linear 4k_pair_linear; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_warp_tiled -s 3 -c 1 -f -o /tmp/${TAG}_${W} python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-other-workloads --workload $W > gpurun_out/${TAG}_${W}_p.log 2>&1
  python scripts/ncu_summary.py /tmp/${TAG}_${W}.ncu-rep ${TAG}_${W} x 0 gpurun_out > /dev/null
  python scripts/ncu_lines.py /tmp/${TAG}_${W}.ncu-rep 80 > gpurun_out/${TAG}_${W}_lines.txt 2>&1
  head -3 gpurun_out/${TAG}_${W}_ncu_summary.txt
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/${TAG}_l.log 2>&1
grep -c k_warp gpurun_out/${TAG}_launches.csv
