#!/bin/bash
# usage: scripts/gpu_pairs.sh TAG [pytest files] -- the given GPU tests (default: fast chain + parity), then the analytic single-pair launches
TAG=$1; shift
mkdir -p gpurun_out
timeout 900 python -m pytest ${@:-tests/test_gpu_fast_chain.py tests/test_gpu_parity.py} -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${TAG}_tests.log
tail -2 gpurun_out/${TAG}_tests.log
for A in "--workload 4k_pair_linear" "--workload 8k_pair_linear" "--workload 8k_rot_poly_linear --pairs 4"; do
    N=$(echo $A | tr -d ' -')
    timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-other-workloads $A > gpurun_out/${TAG}_${N}.json 2>> gpurun_out/${TAG}.err
    python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${TAG}_${N}.json")); print("[$A]", round(d["ms_per_step"]*1e3,2), "us  frac", round(d["roofline"]["frac"],4), d["clocks"].get("sm_mhz"), d["clocks"].get("reasons"))
except Exception as e: print("[$A] failed", e)
PY
done
tail -3 gpurun_out/${TAG}.err
