#!/bin/bash
# usage: scripts/gpu_final_full.sh TAG -- what the driver runs at round end (GPU tests, smoke, default bench line, reference arm)
# + memcheck of the folded-chain tests + ncu --set full of the headline kernel + the launch list of the same command
TAG=$1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${TAG}_tests.log
tail -2 gpurun_out/${TAG}_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -1 gpurun_out/${TAG}_smoke.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
head -c 600 gpurun_out/${TAG}_bench.json; echo; tail -3 gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_reference_arm.json 2>> gpurun_out/${TAG}_bench.err
head -c 400 gpurun_out/${TAG}_reference_arm.json; echo
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_fast_chain.py -x -q -k "matches_the_oracle or per_frame_radius" > gpurun_out/${TAG}_memcheck.log 2>&1
echo "memcheck exit code $?" >> gpurun_out/${TAG}_memcheck.log
tail -4 gpurun_out/${TAG}_memcheck.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_warp_tiled -s 3 -c 1 -f -o /tmp/${TAG}_prof python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-other-workloads > gpurun_out/${TAG}_p.log 2>&1
python scripts/ncu_summary.py /tmp/${TAG}_prof.ncu-rep ${TAG} x 0 gpurun_out > /dev/null
python scripts/ncu_lines.py /tmp/${TAG}_prof.ncu-rep 60 > gpurun_out/${TAG}_lines.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/${TAG}_l.log 2>&1
grep -c k_warp gpurun_out/${TAG}_launches.csv
head -3 gpurun_out/${TAG}_ncu_summary.txt
