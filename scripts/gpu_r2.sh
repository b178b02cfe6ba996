#!/bin/bash
# usage: scripts/gpu_r2.sh TAG [san]  -- round-2 GPU pass: parity tests (incl. forced frames-per-CTA), LSU microbench,
# full default bench line (all workloads, e2e legs, parity of the timed launches), optional compute-sanitizer pass
TAG=$1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q --maxfail=40 -rf 2>&1 | tail -60 > gpurun_out/${TAG}_tests.log
tail -5 gpurun_out/${TAG}_tests.log
[ -x scripts/ubench/lsu_mix ] && timeout 120 scripts/ubench/lsu_mix > gpurun_out/${TAG}_lsu_mix.txt 2>&1
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -5 gpurun_out/${TAG}_bench.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${TAG}_bench.json"))
    print("main", d["ms_per_step"], round(d["roofline"]["frac"],4), "parity", d.get("parity",{}).get("bit_exact"))
    e=d.get("e2e") or {}
    print("e2e", round(e.get("value",0)), "ceiling", round(e.get("copy_ceiling",0)), e.get("copy_gbs"))
    p=d.get("e2e_python_api") or {}
    print("py", round(p.get("value",0)), p.get("vs_c_abi"))
    for k,v in (d.get("other_workloads") or {}).items(): print(k, round(v["ms_per_step"],4), round(v["roofline_frac"],4), v.get("parity_bit_exact"))
    print("cpu", d.get("cpu_baseline",{}).get("value"))
except Exception as ex: print("bench parse failed", ex)
PY
if [ "$2" = "san" ]; then
  timeout 700 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_frameloop.py -x -q -k "fixed_radius and 24 and not True" > gpurun_out/${TAG}_memcheck.log 2>&1
  echo "memcheck rc $?" >> gpurun_out/${TAG}_memcheck.log
  tail -4 gpurun_out/${TAG}_memcheck.log
  timeout 500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_frameloop.py -x -q -k "fixed_radius and 24 and not True and (1-24 or 2-24)" > gpurun_out/${TAG}_racecheck.log 2>&1
  echo "racecheck rc $?" >> gpurun_out/${TAG}_racecheck.log
  tail -4 gpurun_out/${TAG}_racecheck.log
fi
