#!/bin/bash
# usage: scripts/gpu_quick.sh TAG [bench args] -- GPU parity tests + default bench line (+ a second line with the extra args)
TAG=$1; shift
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${TAG}_tests.log
timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -3 gpurun_out/${TAG}_tests.log
python -c "import json,sys; d=json.load(open('gpurun_out/${TAG}_bench.json')); print('default', d['ms_per_step'], d['roofline']['frac'])"
if [ -n "$1" ]; then
  timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e "$@" > gpurun_out/${TAG}_bench2.json 2>> gpurun_out/${TAG}_bench.err
  python -c "import json,sys; d=json.load(open('gpurun_out/${TAG}_bench2.json')); print('$*', d['ms_per_step'], d['roofline']['frac'])"
fi
tail -3 gpurun_out/${TAG}_bench.err
