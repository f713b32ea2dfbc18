#!/bin/bash
# Tuning builds of libskagpu.so (same sources, different compile-time tile shapes) for A/B runs on the GPU box:
#   SKAGPU_LIB=ska.rust_b200/variants/libskagpu_<name>.so python bench.py ...      (or profiles/ab.sh)
# Outputs are git-ignored (*.so) but travel with gpurun.
#   build_variants.sh                       the default set below
#   build_variants.sh <name> <file> "<defines>"   one variant of one source file
set -e
cd "$(dirname "$0")/../ska.rust_b200/csrc"
mkdir -p ../variants
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr"
build() {   # name, file, defines
    $NVCC $FLAGS $3 -c $2.cu -o ../variants/$2_$1.o
    objs=""
    for f in api k1_extract k1q_reads partition merge rowmerge filter distance verify setops merge128 cov sort128 dedup comm; do
        if [ "$f" == "$2" ]; then objs="$objs ../variants/$2_$1.o"; else objs="$objs $f.o"; fi
    done
    $NVCC -gencode arch=compute_100a,code=sm_100a -shared -o ../variants/libskagpu_$1.so $objs
}
if [ $# -gt 0 ]; then
    build "$1" "$2" "$3"
else
    build sc5 partition "-DSKA_SC_MIN_CTAS=5"
    build sc6 partition "-DSKA_SC_MIN_CTAS=6"
    build histnopf partition "-DSKA_HIST_PREFETCH=0"
fi
ls -la ../variants/*.so
