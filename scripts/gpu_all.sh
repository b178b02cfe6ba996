#!/bin/bash
# usage: scripts/gpu_all.sh TAG -- GPU parity tests + all bench workloads (no CPU baseline)
TAG=$1
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${TAG}_tests.log
tail -3 gpurun_out/${TAG}_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 --all-workloads --no-cpu-baseline > gpurun_out/${TAG}_all.json 2> gpurun_out/${TAG}_all.err
python - <<PY
import json; d=json.load(open("gpurun_out/${TAG}_all.json"))
print("default", d["config"]["pairs_per_gpu_per_step"], d["ms_per_step"], round(d["roofline"]["frac"],4), "e2e", round(d["e2e"]["value"]))
for k,v in d["other_workloads"].items(): print(k, round(v["ms_per_step"],4), round(v["roofline_frac"],4))
PY
tail -3 gpurun_out/${TAG}_all.err
