#!/bin/bash
# Builds build/exp/<name>.so: the library with sparc_amp.cu recompiled under extra -D flags (A/B runs, tools/ab_amp.sh)
# usage: tools/build_variant.sh <name> [-DFLAG=VALUE ...]
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p build/exp
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xptxas -v"
nvcc $FLAGS "$@" -c ldpc_sparc_b200/csrc/sparc_amp.cu -o build/exp/$name.o > build/exp/$name.ptxas.log 2>&1
objs=$(ls build/*.o | grep -v sparc_amp.o)
nvcc -shared -o build/exp/$name.so $objs build/exp/$name.o -gencode arch=compute_100a,code=sm_100a
grep -A2 "amp_pair_kernelILi11" build/exp/$name.ptxas.log | grep -E "registers|spill" | head -2
