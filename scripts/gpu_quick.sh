#!/bin/bash
# usage: scripts/gpu_quick.sh TAG -- GPU parity tests + default bench line, each under its own timeout; outputs in gpurun_out/TAG_*
TAG=$1
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${TAG}_tests.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
cat gpurun_out/${TAG}_tests.log
cat gpurun_out/${TAG}_bench.json
tail -5 gpurun_out/${TAG}_bench.err
