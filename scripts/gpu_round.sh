#!/bin/bash
# usage: scripts_gpu_round.sh TAG  -- tests, bench (default + pairs 32), ncu of the top kernel; outputs in gpurun_out/TAG_*
TAG=$1
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${TAG}_tests.log
python bench.py --steps 10 --warmup 3 --all-workloads --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python bench.py --steps 10 --warmup 3 --pairs 16 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_bench_p16.json 2>> gpurun_out/${TAG}_bench.err
ncu --set full --clock-control none --import-source on -k regex:k_warp_tiled -s 3 -c 1 -o gpurun_out/${TAG}_prof python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/${TAG}_p.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/${TAG}_l.log 2>&1
