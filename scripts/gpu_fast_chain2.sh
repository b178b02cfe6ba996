#!/bin/bash
# usage: scripts/gpu_fast_chain2.sh TAG -- GPU suite, then launches where the chain weighs most (one pair, 8 pairs)
TAG=$1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${TAG}_tests.log
tail -2 gpurun_out/${TAG}_tests.log
for A in "--workload 4k_pair_linear" "--workload 8k_pair_linear" "--workload 8k_rot_poly_linear --pairs 8" "--workload 8k_rot_poly_linear"; do
    N=$(echo $A | tr -d ' -')
    timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-other-workloads $A > gpurun_out/${TAG}_${N}.json 2>> gpurun_out/${TAG}.err
    python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${TAG}_${N}.json")); print("[$A]", round(d["ms_per_step"]*1e3,2), "us  frac", round(d["roofline"]["frac"],4), d.get("parity",{}).get("bit_exact"), d["clocks"].get("sm_mhz"), d["clocks"].get("reasons"))
except Exception as e: print("[$A] failed", e)
PY
done
tail -3 gpurun_out/${TAG}.err
