#!/bin/bash
# usage: scripts/gpu_san.sh TAG -- compute-sanitizer memcheck + racecheck over the long-frame-loop tests of the tiled kernel
TAG=$1
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_frameloop.py -x -q -k "(fixed_radius and 24 and not True) or (per_frame_radius and runs and True) or (every_border and border1) or (lut_sources and lut_packed)" > gpurun_out/${TAG}_memcheck.log 2>&1
echo "memcheck exit code $?" >> gpurun_out/${TAG}_memcheck.log
tail -6 gpurun_out/${TAG}_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_frameloop.py -x -q -k "(fixed_radius and 24 and not True and (1-24 or 2-24)) or (per_frame_radius and runs and False-1)" > gpurun_out/${TAG}_racecheck.log 2>&1
echo "racecheck exit code $?" >> gpurun_out/${TAG}_racecheck.log
tail -6 gpurun_out/${TAG}_racecheck.log
