#!/bin/bash
# Builds variants of libpv_b200.so that differ in -D switches of ONE translation unit, for A/B runs on the GPU box:
#   tools/ab_variants.sh pv_fused_corrected_kernels base: nofold:-DPV_EXP_NO_WINFOLD ...
#   gpurun -- 'for v in base nofold; do PV_B200_LIB=phase-vocoder_b200/build/variants/libpv_b200_$v.so python bench.py ...; done'
# (build/ is git-ignored but travels to the GPU box.)
set -e
cd "$(dirname "$0")/../phase-vocoder_b200"
TU=$1; shift
make > /dev/null
mkdir -p build/variants
for spec in "$@"; do
    name=${spec%%:*}; flags=${spec#*:}; [ "$flags" = "$spec" ] && flags=""
    ( /usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Xptxas -v $flags \
        -c -o build/variants/${TU}_$name.o csrc/$TU.cu 2> build/variants/${TU}_$name.log
      objs=$(ls build/*.o | grep -v "build/$TU.o")
      /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/variants/libpv_b200_$name.so $objs build/variants/${TU}_$name.o
      echo "$name: $(grep -A3 'corrected_fused_kernelILi11ELi4\|compat_fused_kernelILi11' build/variants/${TU}_$name.log | grep -E 'Used|spill' | tr '\n' ' ')" ) &
done
wait
