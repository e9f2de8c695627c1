#!/bin/bash
# development aid: kernel time per pack (chained launches, K=50) for library variants, interleaved reps
# usage: ab_clip.sh out_file variant...   ("main" = in-tree)
out=$1; shift
for rep in 1 2 3; do
  for v in "$@"; do
    if [ "$v" == "main" ]; then unset V2E_LIB; else export V2E_LIB=$PWD/variants/lib_$v.so; fi
    python scripts/clip_probe.py philox fixpoint 50 2>&1 | awk -v v=$v '/rep [23]/ {print v, $0}' >> $out
  done
done
