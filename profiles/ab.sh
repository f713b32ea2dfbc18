#!/bin/bash
# A/B runs of library variants on the GPU box: profiles/ab.sh <tag> [variant ...]  ("" = the in-tree build)
# prints the stage timers of the device-resident arm and the step time per variant
tag=$1; shift
mkdir -p gpurun_out
for v in "$@"; do
  lib=""; [ "$v" != "base" ] && lib="ska.rust_b200/variants/libskagpu_$v.so"
  SKAGPU_LIB=$lib python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ab_${tag}_$v.json 2> gpurun_out/ab_${tag}_$v.err
  echo "== $v: $(grep -o 'device warm-up done.*' gpurun_out/ab_${tag}_$v.err | cut -c1-600)"
  grep -o 'device arm:.*' gpurun_out/ab_${tag}_$v.err
done
