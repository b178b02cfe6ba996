#!/bin/bash
# usage: scripts/gpu_prof.sh TAG -- GPU parity tests, default bench line, one ncu --set full capture of the warp kernel
TAG=$1
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${TAG}_tests.log
timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_warp_tiled -s 3 -c 1 -o gpurun_out/${TAG}_prof python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline $2 > gpurun_out/${TAG}_p.log 2>&1
tail -4 gpurun_out/${TAG}_tests.log
cat gpurun_out/${TAG}_bench.json
tail -3 gpurun_out/${TAG}_bench.err
