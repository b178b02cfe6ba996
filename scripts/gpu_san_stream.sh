#!/bin/bash
# usage: scripts/gpu_san_stream.sh TAG -- compute-sanitizer memcheck + racecheck over the tile-streaming kernel's tests
TAG=$1
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_stream.py -x -q -k "(matches_oracle and 1-3) or (matches_oracle and 2-3 and True) or more_unstaged or (borders and colour)" > gpurun_out/${TAG}_memcheck.log 2>&1
echo "memcheck exit code $?" >> gpurun_out/${TAG}_memcheck.log
tail -6 gpurun_out/${TAG}_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_stream.py -x -q -k "(matches_oracle and False-1-1-3) or (matches_oracle and True-2-2-3) or (borders and colour-1)" > gpurun_out/${TAG}_racecheck.log 2>&1
echo "racecheck exit code $?" >> gpurun_out/${TAG}_racecheck.log
tail -6 gpurun_out/${TAG}_racecheck.log
