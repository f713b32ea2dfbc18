#!/bin/bash
# A/B runs of library variants on the GPU box: profiles/ab.sh <tag> <spec> ...   spec = variant[:ENV=VALUE[,ENV=VALUE]]
# ("base" = the in-tree build; variants come from profiles/build_variants.sh).  Prints the stage timers of the
# device-resident arm and the step time per spec.
tag=$1; shift
mkdir -p gpurun_out
for spec in "$@"; do
  v=${spec%%:*}; envs=""; [ "$spec" != "$v" ] && envs=${spec#*:}
  lib=""; [ "$v" != "base" ] && lib="ska.rust_b200/variants/libskagpu_$v.so"
  name=$(echo "$spec" | tr ':=,' '___')
  env SKAGPU_LIB=$lib $(echo "$envs" | tr ',' ' ') python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e $AB_ARGS > gpurun_out/ab_${tag}_$name.json 2> gpurun_out/ab_${tag}_$name.err
  echo "== $spec: $(grep -o 'device warm-up done.*' gpurun_out/ab_${tag}_$name.err | cut -c1-600)"
  grep -o 'device arm:.*' gpurun_out/ab_${tag}_$name.err
done
