#!/bin/bash
# usage: scripts/gpu_round.sh TAG  -- GPU tests, full bench line (all workloads, e2e, CPU baseline), 32-pair line,
# reference arm, ncu --set full of the top kernel and the launch list of the same bench command; outputs in gpurun_out/TAG_*
TAG=$1
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${TAG}_tests.log
timeout 900 python bench.py --all-workloads > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
timeout 200 python bench.py --steps 10 --warmup 3 --pairs 32 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_bench_p32.json 2>> gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_reference_arm.json 2>> gpurun_out/${TAG}_bench.err
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_warp_tiled -s 3 -c 1 -o gpurun_out/${TAG}_prof python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/${TAG}_p.log 2>&1
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --all-workloads > gpurun_out/${TAG}_l.log 2>&1
tail -3 gpurun_out/${TAG}_tests.log
head -c 1500 gpurun_out/${TAG}_bench.json; echo
tail -3 gpurun_out/${TAG}_bench.err
for P in 16 128; do
  timeout 200 python bench.py --steps 10 --warmup 3 --pairs $P --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_bench_p${P}.json 2>> gpurun_out/${TAG}_bench.err
done
python - <<PY
import json
for P in (16, 32, 128):
    try:
        d = json.load(open("gpurun_out/${TAG}_bench_p%d.json" % P)); print("pairs", P, d["ms_per_step"], round(d["roofline"]["frac"], 4))
    except Exception as e: print("pairs", P, "failed", e)
PY
