#!/bin/bash
# Builds ldpc_sparc_b200/libldpcsparc_b200.so (sm_100a) and the CPU oracle.
# Translation units compile in parallel; objects are cached under build/.
set -e
cd "$(dirname "$0")"
SRC=ldpc_sparc_b200/csrc
OUT=ldpc_sparc_b200/libldpcsparc_b200.so
mkdir -p build
NVCC=${NVCC:-nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xptxas -v"
pids=()
objs=()
for f in $SRC/*.cu; do
  o=build/$(basename "$f" .cu).o
  objs+=("$o")
  if [ ! -f "$o" ] || [ -n "$(find $SRC include -newer "$o" \( -name '*.cu' -o -name '*.cuh' -o -name '*.h' \) | head -1)" ]; then
    ( $NVCC $FLAGS -c "$f" -o "$o" > "build/$(basename "$f" .cu).ptxas.log" 2>&1 || { cat "build/$(basename "$f" .cu).ptxas.log" | grep -v "^ptxas info" >&2; exit 1; } ) &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait "$p"; done
$NVCC -shared -o "$OUT" "${objs[@]}" -gencode arch=compute_100a,code=sm_100a
make -s -C oracle
echo "built $OUT"
