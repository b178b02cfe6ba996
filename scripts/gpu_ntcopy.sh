#!/bin/bash
# usage: scripts/gpu_ntcopy.sh TAG -- A/B of the copy threads' streaming stores (VR180_NT_COPY=0 / 1): the Python-API e2e leg on pageable arrays
TAG=$1
mkdir -p gpurun_out
for round in 1 2; do
for NT in 0 1; do
  VR180_NT_COPY=$NT timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-other-workloads > gpurun_out/${TAG}_nt${NT}_${round}.json 2>> gpurun_out/${TAG}.err
  python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_nt${NT}_${round}.json")); p=d["e2e_python_api"]
print("NT=$NT round $round: python pageable", round(p["value"]), "Mpix/s  vs C-ABI", round(p["vs_c_abi"],3), " pinned", round(p.get("pinned_inputs_per_gpu_value",0)), " C-ABI e2e", round(d["e2e"]["value"]))
PY
done
done
tail -2 gpurun_out/${TAG}.err
