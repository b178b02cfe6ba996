#!/bin/bash
# usage: scripts/gpu_exp.sh TAG "args1" "args2" ...  -- one short bench line (10 steps, no e2e / CPU legs / other workloads) per argument string
TAG=$1; shift
mkdir -p gpurun_out
i=0
for A in "$@"; do
  timeout 300 python bench.py --steps ${STEPS:-10} --warmup 3 --no-e2e --no-cpu-baseline --no-other-workloads $A > gpurun_out/${TAG}_exp$i.json 2>> gpurun_out/${TAG}_exp.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${TAG}_exp$i.json")); print("[$A]", round(d["ms_per_step"],4), "ms  frac", round(d["roofline"]["frac"],4), " clocks", d["clocks"].get("sm_mhz"), d["clocks"].get("reasons"))
except Exception as e: print("[$A] failed", e)
PY
  i=$((i+1))
done
tail -3 gpurun_out/${TAG}_exp.err
