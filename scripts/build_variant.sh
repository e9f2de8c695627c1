#!/bin/bash
# development aid: build a library variant side by side:
#   build_variant.sh NAME [file.cu ...] -- "-DFLAG ..."   (default file: v2e_fixpoint_spec.cu)
# -> variants/lib_NAME.so  (select it at run time with V2E_LIB=variants/lib_NAME.so; the directory
# travels to the GPU box, the .so files are git-ignored)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
files=()
while [ $# -gt 0 ] && [ "$1" != "--" ]; do files+=("$1"); shift; done
[ "$1" == "--" ] && shift
[ ${#files[@]} -eq 0 ] && files=(v2e_fixpoint_spec.cu)
out=variants/obj_$name; mkdir -p $out
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --fmad=false -Xcompiler -fPIC -Xptxas -v -DV2E_FX_DEV"
objs=()
for o in v2e2v_b200/_build/*.o; do
  b=$(basename $o .o)
  skip=0; for f in "${files[@]}"; do [ "$b" == "$f" ] && skip=1; done
  [ $skip -eq 0 ] && objs+=($o)
done
for f in "${files[@]}"; do
  nvcc $FLAGS $@ -c v2e2v_b200/csrc/$f -o $out/$f.o 2> $out/$f.ptxas.log &
  objs+=($out/$f.o)
done
wait
nvcc -shared -o variants/lib_$name.so "${objs[@]}" -gencode arch=compute_100a,code=sm_100a -cudart static
echo variants/lib_$name.so
