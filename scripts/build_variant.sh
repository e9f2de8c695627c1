#!/bin/bash
# development aid: build a library variant side by side:  build_variant.sh NAME "-DV2E_FX_THREADS=768"
# -> v2e2v_b200/_build/lib_NAME.so  (select it at run time with V2E_LIB=...)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
out=v2e2v_b200/_build/var_$name; mkdir -p $out
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --fmad=false -Xcompiler -fPIC -Xptxas -v"
for f in v2e_api v2e_stepwise v2e_persistent v2e_runahead v2e_voxel; do
  cp v2e2v_b200/_build/$f.cu.o $out/ 2>/dev/null || nvcc $FLAGS -c v2e2v_b200/csrc/$f.cu -o $out/$f.cu.o
done
nvcc $FLAGS "$@" -c v2e2v_b200/csrc/v2e_fixpoint.cu -o $out/v2e_fixpoint.cu.o 2> $out/ptxas.log
cp v2e2v_b200/_build/host_scalars.c.o $out/
nvcc -shared -o v2e2v_b200/_build/lib_$name.so $out/*.o -gencode arch=compute_100a,code=sm_100a -cudart static
echo v2e2v_b200/_build/lib_$name.so
