#!/bin/bash
# Tuning builds of libskagpu.so (same sources, different compile-time tile shapes) for A/B runs on the GPU box:
#   SKAGPU_LIB=ska.rust_b200/variants/libskagpu_<name>.so python bench.py ...
# Outputs are git-ignored (*.so) but travel with gpurun.
set -e
cd "$(dirname "$0")/../ska.rust_b200/csrc"
mkdir -p ../variants
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr"
build() {   # name, file, defines
    $NVCC $FLAGS $3 -c $2.cu -o ../variants/$2_$1.o
    objs=""
    for f in api k1_extract k1q_reads partition merge rowmerge filter distance verify setops merge128 cov sort128; do
        if [ "$f" == "$2" ]; then objs="$objs ../variants/$2_$1.o"; else objs="$objs $f.o"; fi
    done
    $NVCC -gencode arch=compute_100a,code=sm_100a -shared -o ../variants/libskagpu_$1.so $objs
}
build sc256_3 partition "-DSKA_SC_THREADS=256 -DSKA_SC_MIN_CTAS=3"
build sc512_2 partition "-DSKA_SC_THREADS=512 -DSKA_SC_MIN_CTAS=2"
build sc512_3 partition "-DSKA_SC_THREADS=512 -DSKA_SC_MIN_CTAS=3"
ls -la ../variants/*.so
