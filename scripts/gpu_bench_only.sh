#!/bin/bash
# usage: scripts/gpu_bench_only.sh -- default bench line twice (10 steps each), no tests
for i in 1 2; do
  timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], round(d['roofline']['frac'],4))"
done
