#!/bin/bash
# usage: scripts/gpu_round_end.sh TAG -- what the driver runs at round end: GPU tests, smoke, the default bench line
TAG=$1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${TAG}_tests.log
tail -2 gpurun_out/${TAG}_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -1 gpurun_out/${TAG}_smoke.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -3 gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.load(open('gpurun_out/${TAG}_bench.json'))
print(d['ms_per_step'], d['roofline']['frac'], d['clocks'], d['parity']['bit_exact'])
print('e2e', d['e2e']['value'], d['e2e'].get('frac_of_ceiling'), 'py', d['e2e_python_api']['value'], 'cpu', d['cpu_baseline']['value'])
for n,w in d['other_workloads'].items(): print(f"{n:26s} {w['ms_per_step']*1e3:9.1f} us frac {w['roofline_frac']:.4f} parity {w.get('parity_bit_exact')}")
PY
