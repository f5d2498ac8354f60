#!/bin/bash
# Runs on the GPU box: rebuilds libuvcgpu.so with different compile-time shapes (EXTRA_DEFS) and prints the per-kernel times of one sub-batch of
# configs[1] for each. usage: tools/gpu_variants.sh "<defs of variant 1>" "<defs of variant 2>" ...   (an empty string = the default build)
mkdir -p gpurun_out
for V in "$@"; do
  touch uvc_b200/csrc/engine.cpp
  make -C uvc_b200/csrc EXTRA_DEFS="$V" "$PWD/uvc_b200/lib/libuvcgpu.so" -j4 > gpurun_out/variant_build.log 2>&1 || { echo "build failed"; tail -5 gpurun_out/variant_build.log; continue; }
  echo "=== variant [$V]"; grep -E "uvc_k2_bias|uvc_k3b|uvc_k4_family_consensus9" -A3 uvc_b200/lib/engine.ptxas.log | grep -E "Used|spill" | head -8
  python bench.py --config c2 --scale 0.05 --steps 5 --warmup 3 --sub-batches 1 --contexts 1 --skip-cpu-baseline --skip-pipeline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('ms/step %.2f' % d['ms_per_step'], {k:round(v,2) for k,v in d['stage_ms_per_step'].items() if v>0.3})"
done
touch uvc_b200/csrc/engine.cpp; make -C uvc_b200/csrc "$PWD/uvc_b200/lib/libuvcgpu.so" -j4 > /dev/null 2>&1
