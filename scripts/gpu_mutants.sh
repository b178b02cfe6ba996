#!/bin/bash
# usage: scripts/gpu_mutants.sh TAG -- mutation check of the frame-loop tests: the library is
# replaced by a build with ONE line of csrc/tiled.cu changed (mutants/*.so, built in the container by the recipe in
# profiles/README.md); the tests must FAIL on each mutant and pass on the real build.
#   m1: the sampling warps' stage-ring phase flip removed (`if (++st == S) { st = 0; ph ^= 1u; }` -> `{ st = 0; }`)
#   m2: the per-stage rectangle origin of per-frame-radius launches not published (`s_org[p_stage] = make_int2(bx0, ry0)` -> (0, 0))
TAG=$1
mkdir -p gpurun_out
cp vr180-convert_b200/libvr180_b200.so /tmp/real.so
for m in m1 m2; do
  cp mutants/$m.so vr180-convert_b200/libvr180_b200.so
  timeout 600 python -m pytest tests/test_gpu_frameloop.py -q -x 2>&1 | tail -4 > gpurun_out/${TAG}_$m.log
  echo "== $m: $(tail -1 gpurun_out/${TAG}_$m.log)"
done
cp /tmp/real.so vr180-convert_b200/libvr180_b200.so
timeout 600 python -m pytest tests/test_gpu_frameloop.py -q -x 2>&1 | tail -2 > gpurun_out/${TAG}_real.log
echo "== real build: $(tail -1 gpurun_out/${TAG}_real.log)"
